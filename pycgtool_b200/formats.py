"""Minimal GRO / XTC readers and a GRO writer.

The reference delegates file I/O to mdtraj (``pycgtool/frame.py:41-63,231-244``),
which is not available here; this module covers what the hot path's callers
need to run the reference's own fixtures: GROMACS ``.gro`` (read / write, with
mdtraj's water renaming) and ``.xtc`` (read).  I/O is host-side Python and is
not part of the accelerated path.
"""

import struct

import numpy as np

from .topology import Topology, Trajectory

# mdtraj renames residues / atoms through its PDB tables when it builds a
# topology (see reference tests/data/water.gro:3-4 and mapping.py:288-317).
# Only the water entry is reproduced; everything else keeps its file name.
WATER_RESIDUES = ("HOH", "H2O", "OH2", "SOL", "WAT", "TIP", "TIP2", "TIP3", "TIP4")
RESIDUE_NAME_REPLACEMENTS = {name: "HOH" for name in WATER_RESIDUES}
ATOM_NAME_REPLACEMENTS = {
    "HOH": {"O": "O", "OW": "O", "OH2": "O", "H1": "H1", "HW1": "H1", "H2": "H2", "HW2": "H2"},
}


def read_gro(path):
    """Single-frame .gro -> ``Trajectory`` (nm, float32)."""
    with open(path) as handle:
        text = handle.read().split("\n")
    natoms = int(text[1])
    body = text[2:2 + natoms]
    residues, last = [], None
    xyz = np.empty((1, natoms, 3), dtype=np.float32)
    if natoms:
        dots = [i for i, ch in enumerate(body[0]) if ch == "." and i >= 20]
        width = dots[1] - dots[0] if len(dots) > 1 else 8
    for i, line in enumerate(body):
        key = (line[0:5], line[5:10])
        if key != last:
            resname = RESIDUE_NAME_REPLACEMENTS.get(line[5:10].strip(), line[5:10].strip())
            residues.append((resname, []))
            last = key
        resname = residues[-1][0]
        atom = line[10:15].strip()
        residues[-1][1].append(ATOM_NAME_REPLACEMENTS.get(resname, {}).get(atom, atom))
        xyz[0, i] = [float(line[20 + k * width:20 + (k + 1) * width]) for k in range(3)]
    fields = [float(tok) for tok in text[2 + natoms].split()]
    box = np.zeros((1, 3, 3), dtype=np.float32)
    if len(fields) >= 3:
        box[0, 0, 0], box[0, 1, 1], box[0, 2, 2] = fields[:3]
    if len(fields) == 9:
        box[0, 0, 1], box[0, 0, 2], box[0, 1, 0], box[0, 1, 2], box[0, 2, 0], box[0, 2, 1] = fields[3:]
    top = Topology.from_residue_lists(residues)
    return Trajectory(xyz, top, np.zeros(1, dtype=np.float32), box if box.any() else None)


def write_gro(path, traj, frame=0):
    """Write one frame in mdtraj's .gro dialect (zero-based numbering for
    topologies built in memory, 3 decimals, always nine box fields)."""
    top = traj.topology
    xyz = traj.xyz[frame]
    lines = ["Generated with MDTraj, t= %s" % float(traj.time[frame]), " %d" % top.n_atoms]
    res_of_atom = np.repeat(np.arange(top.n_residues), np.diff(top.res_first))
    for i in range(top.n_atoms):
        r = int(res_of_atom[i])
        lines.append("%5d%-5s%5s%5d%8.3f%8.3f%8.3f" % (
            r % 100000, top.res_names[top.res_name_id[r]], top.atom_names[top.atom_name_id[i]],
            i % 100000, xyz[i, 0], xyz[i, 1], xyz[i, 2]))
    v = traj.unitcell_vectors
    b = np.zeros((3, 3)) if v is None else v[frame]
    lines.append(("%10.5f" * 9) % (b[0, 0], b[1, 1], b[2, 2], b[0, 1], b[0, 2], b[1, 0], b[1, 2], b[2, 0], b[2, 1]))
    with open(path, "w") as handle:
        handle.write("\n".join(lines) + "\n")


# ------------------------------------------------------------------------------- XTC
_FIRST = 9
_MAGIC = (0,) * 9 + (
    8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
    32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
    524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304,
    5284491, 6658042, 8388607, 10568983, 13316085, 16777216)


def _unpack_frame(natoms, lo, hi, smallidx, payload):
    """xdrfile ``xdr3dfcoord`` integer decoder for one frame -> (natoms, 3) int32."""
    size = [hi[d] - lo[d] + 1 for d in range(3)]
    wide = any(s > 0xFFFFFF for s in size)
    nbits_each = [s.bit_length() for s in size]
    nbits_all = (size[0] * size[1] * size[2]).bit_length()
    stream = int.from_bytes(payload, "big")
    remaining = 8 * len(payload)

    def bits(n):
        nonlocal remaining
        remaining -= n
        return (stream >> remaining) & ((1 << n) - 1)

    def triple(n, radix):
        # bytes arrive least-significant first; the last chunk may be partial
        acc, shift = 0, 0
        while n > 8:
            acc |= bits(8) << shift
            shift += 8
            n -= 8
        if n:
            acc |= bits(n) << shift
        acc, z = divmod(acc, radix[2])
        x, y = divmod(acc, radix[1])
        return x, y, z

    half = _MAGIC[smallidx] // 2
    below = _MAGIC[max(_FIRST, smallidx - 1)] // 2
    out = np.empty((natoms + 8, 3), dtype=np.int64)
    n, run = 0, 0
    while n < natoms:
        if wide:
            cur = [bits(nbits_each[d]) + lo[d] for d in range(3)]
        else:
            t = triple(nbits_all, size)
            cur = [t[d] + lo[d] for d in range(3)]
        change = 0
        if bits(1):
            run = bits(5)
            change = run % 3
            run -= change
            change -= 1
        if run > 0:
            small_radix = (_MAGIC[smallidx],) * 3
            anchor = cur
            for k in range(0, run, 3):
                t = triple(smallidx, small_radix)
                nxt = [t[d] + anchor[d] - half for d in range(3)]
                if k == 0:
                    out[n] = nxt        # water trick: the second atom is stored first
                    out[n + 1] = anchor
                    n += 2
                else:
                    out[n] = nxt
                    n += 1
                anchor = nxt
        else:
            out[n] = cur
            n += 1
        smallidx += change
        if change < 0:
            half = below
            below = _MAGIC[smallidx - 1] // 2 if smallidx > _FIRST else 0
        elif change > 0:
            below = half
            half = _MAGIC[smallidx] // 2
    return out[:natoms].astype(np.int32)


def read_xtc(path, topology=None):
    """.xtc -> ``Trajectory``.  Coordinates are ``float32(int) * float32(1/precision)``."""
    with open(path, "rb") as handle:
        blob = handle.read()
    pos, frames, times, boxes = 0, [], [], []
    while pos < len(blob):
        magic, natoms, _step, time = struct.unpack_from(">iiif", blob, pos)
        if magic != 1995:
            raise OSError("XTC magic number mismatch: format is not supported")
        boxes.append(np.frombuffer(blob, ">f4", 9, pos + 16).astype(np.float32).reshape(3, 3))
        (count,) = struct.unpack_from(">i", blob, pos + 52)
        pos += 56
        if count <= 9:
            coords = np.frombuffer(blob, ">f4", 3 * count, pos).astype(np.float32)
            pos += 12 * count
        else:
            precision, = struct.unpack_from(">f", blob, pos)
            lo = struct.unpack_from(">3i", blob, pos + 4)
            hi = struct.unpack_from(">3i", blob, pos + 16)
            smallidx, nbytes = struct.unpack_from(">2i", blob, pos + 28)
            pos += 36
            ints = _unpack_frame(count, lo, hi, smallidx, blob[pos:pos + nbytes])
            pos += -(-nbytes // 4) * 4
            coords = ints.astype(np.float32).reshape(-1) * (np.float32(1.0) / np.float32(precision))
        frames.append(coords.reshape(count, 3))
        times.append(time)
    xyz = np.stack(frames).astype(np.float32)
    if topology is not None and topology.n_atoms != xyz.shape[1]:
        raise ValueError("Number of atoms does not match between topology and trajectory files")
    box = np.stack(boxes)
    return Trajectory(xyz, topology, np.array(times, dtype=np.float32), box if box.any() else None)


def load(path, top=None):
    """Dispatch on extension, like ``mdtraj.load`` for the supported formats."""
    suffix = str(path).lower().rsplit(".", 1)[-1]
    if suffix == "gro":
        return read_gro(path)
    if suffix == "xtc":
        return read_xtc(path, top)
    raise OSError(f"'{path}': format is not supported")

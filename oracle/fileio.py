"""Oracle file readers: CFG (.map/.bnd/.itp), GRO, XTC.  TEST INFRASTRUCTURE ONLY.

* ``read_cfg``  follows ``pycgtool/parsers/cfg.py:45-85`` (comments, sections,
  ``#include``, duplicate-section and no-section errors).
* ``read_itp``  follows ``pycgtool/parsers/itp.py:18-62``.
* ``read_gro``  restates what ``mdtraj.load`` yields for a .gro file as consumed
  by ``pycgtool/frame.py:41-63,138-159`` (fixed columns, nm, box line of 3 or
  9 fields, PDB-table renaming of water -- see ``tests/data/water.gro:3-4``).
* ``read_xtc``  restates the public xdrfile ``xdr3dfcoord`` decoder that
  mdtraj 1.9.7 wraps (``pycgtool/frame.py:63``); spec in SURVEY.md App. C.1.
"""

import collections
import pathlib
import struct

import numpy as np


# --------------------------------------------------------------------------- CFG
def _strip(line):
    return line.split(";")[0].strip()


def read_cfg(path):
    """Parse a CFG file into ``OrderedDict[section -> list[tuple[str]]]``.

    Reference: ``parsers/cfg.py:45-85``.  ``#include "f"`` merges the sections
    of ``f`` (resolved next to the including file); a section that appears
    twice in one file raises ``KeyError``; tokens before any header raise
    ``KeyError``.
    """
    path = pathlib.Path(path)
    out = collections.OrderedDict()
    current = None
    with open(path) as handle:
        for raw in handle:
            line = _strip(raw)
            if line.startswith("#include"):
                inc = line.split()[1].strip('"')
                out.update(read_cfg(path.parent / inc))
                continue
            if not line:
                continue
            if line.startswith("["):
                current = line.strip("[ ]")
                if current in out:
                    raise KeyError(f"Section '{current}' appears twice in file '{path}'.")
                out[current] = []
                continue
            if current not in out:
                raise KeyError(f"File '{path}' contains no '[]' section headers.")
            out[current].append(tuple(line.split()))
    return out


def read_itp(path):
    """Parse a GROMACS .itp/.top into ``{molecule: {section: [tokens]}}``.

    Reference: ``parsers/itp.py:18-62`` (only atoms/bonds/angles/dihedrals are
    kept; ``#include`` is followed; a molecule defined twice is an error).
    """
    path = pathlib.Path(path)
    keep = ("atoms", "bonds", "angles", "dihedrals")
    out = collections.OrderedDict()
    section, mol = None, None
    with open(path) as handle:
        for raw in handle:
            line = _strip(raw)
            if line.startswith("#include"):
                inc = line.split()[1].strip('"')
                out.update(read_itp(path.parent / inc))
                continue
            if not line:
                continue
            if line.startswith("["):
                section = line.strip("[ ]")
                if section in keep:
                    out[mol].setdefault(section, [])
                continue
            toks = tuple(line.split())
            if section == "moleculetype":
                mol = toks[0]
                if mol in out:
                    raise KeyError(f"Section '{mol}' appears twice in file '{path}'.")
                out[mol] = collections.OrderedDict()
            elif section in keep:
                out[mol][section].append(toks)
            elif section is None:
                raise KeyError(f"File '{path}' contains no '[]' section headers.")
    return out


# --------------------------------------------------------------------------- GRO
# Subset of mdtraj's pdbNames.xml replacement tables (mdtraj 1.9.7,
# formats/pdb/data/pdbNames.xml) -- the water entry, which is the only one the
# reference fixtures exercise (tests/data/water.gro:3-4, tests/test_mapping.py:55-65).
RESIDUE_RENAMES = {n: "HOH" for n in ("HOH", "H2O", "OH2", "SOL", "WAT", "TIP", "TIP2", "TIP3", "TIP4")}
ATOM_RENAMES = {"HOH": {"O": "O", "OW": "O", "OH2": "O", "H1": "H1", "HW1": "H1", "H2": "H2", "HW2": "H2"}}


def rename_residue(name):
    return RESIDUE_RENAMES.get(name, name)


def rename_atom(resname, atomname):
    return ATOM_RENAMES.get(resname, {}).get(atomname, atomname)


def read_gro(path):
    """Read a single-frame GROMACS .gro file.

    Returns ``(residues, xyz, box_vectors)`` with ``residues`` a list of
    ``(resname, [atomname, ...])`` in file order, ``xyz`` float32 ``(1, N, 3)``
    in nm and ``box_vectors`` float32 ``(1, 3, 3)`` (rows a, b, c) or ``None``
    when the box line is all zeros (``tests/test_frame.py:116-118``).

    A new residue starts whenever resSeq or resName changes, as in mdtraj's
    GRO reader.  Box line of 9 fields is ``v1x v2y v3z v1y v1z v2x v2z v3x v3y``.
    """
    with open(path) as handle:
        lines = handle.read().split("\n")
    natoms = int(lines[1])
    residues = []
    xyz = np.zeros((1, natoms, 3), dtype=np.float32)
    last = None
    # Column width of the coordinate fields is derived from the spacing of the
    # decimal points, as the format allows variable precision.
    first = lines[2]
    dots = [i for i, ch in enumerate(first) if ch == "." and i >= 20]
    width = dots[1] - dots[0] if len(dots) > 1 else 8
    for i in range(natoms):
        line = lines[2 + i]
        resseq = int(line[0:5])
        resname = line[5:10].strip()
        atomname = line[10:15].strip()
        key = (resseq, resname)
        if key != last:
            residues.append((rename_residue(resname), []))
            last = key
        residues[-1][1].append(rename_atom(residues[-1][0], atomname))
        for k in range(3):
            xyz[0, i, k] = float(line[20 + k * width: 20 + (k + 1) * width])
    box_fields = [float(tok) for tok in lines[2 + natoms].split()]
    box = np.zeros((1, 3, 3), dtype=np.float32)
    if len(box_fields) >= 3:
        box[0, 0, 0], box[0, 1, 1], box[0, 2, 2] = box_fields[:3]
    if len(box_fields) == 9:
        (box[0, 0, 1], box[0, 0, 2], box[0, 1, 0],
         box[0, 1, 2], box[0, 2, 0], box[0, 2, 1]) = box_fields[3:]
    if not box.any():
        box = None
    return residues, xyz, box


# --------------------------------------------------------------------------- XTC
_FIRSTIDX = 9
_MAGICINTS = [0] * 9 + [
    8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
    32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
    524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304,
    5284491, 6658042, 8388607, 10568983, 13316085, 16777216]


class _Bits:
    """MSB-first bit reader over the opaque block of one XTC frame."""

    def __init__(self, data):
        self.big = int.from_bytes(data, "big")
        self.left = 8 * len(data)

    def take(self, nbits):
        self.left -= nbits
        return (self.big >> self.left) & ((1 << nbits) - 1)

    def take_ints(self, nbits, sizes):
        """xdrfile ``receiveints``: a little-endian base-256 number, read in
        8-bit pieces (last piece may be shorter), unpacked by mixed radix."""
        value, shift = 0, 0
        while nbits > 8:
            value |= self.take(8) << shift
            shift += 8
            nbits -= 8
        if nbits > 0:
            value |= self.take(nbits) << shift
        value, n2 = divmod(value, sizes[2])
        n0, n1 = divmod(value, sizes[1])
        return [n0, n1, n2]


def _decode_coords(natoms, minint, maxint, smallidx, data):
    sizeint = [maxint[k] - minint[k] + 1 for k in range(3)]
    if any(s > 0xFFFFFF for s in sizeint):
        bitsizeint = [s.bit_length() for s in sizeint]
        bitsize = 0
    else:
        bitsizeint = None
        bitsize = (sizeint[0] * sizeint[1] * sizeint[2]).bit_length()
    smaller = _MAGICINTS[max(_FIRSTIDX, smallidx - 1)] // 2
    smallnum = _MAGICINTS[smallidx] // 2
    sizesmall = [_MAGICINTS[smallidx]] * 3
    bits = _Bits(data)
    out = []
    run = 0
    while len(out) < natoms:
        if bitsize == 0:
            this = [bits.take(bitsizeint[k]) for k in range(3)]
        else:
            this = bits.take_ints(bitsize, sizeint)
        this = [this[k] + minint[k] for k in range(3)]
        prev = this
        is_smaller = 0
        if bits.take(1):
            run = bits.take(5)
            is_smaller = run % 3
            run -= is_smaller
            is_smaller -= 1
        if run > 0:
            for k in range(0, run, 3):
                small = bits.take_ints(smallidx, sizesmall)
                cur = [small[d] + prev[d] - smallnum for d in range(3)]
                if k == 0:
                    # first small atom is emitted before the full-precision one
                    out.append(cur)
                    out.append(prev)
                    # after the swap the running reference stays the swapped-in atom
                    prev = cur
                else:
                    out.append(cur)
                    prev = cur
        else:
            out.append(this)
        smallidx += is_smaller
        if is_smaller < 0:
            smallnum = smaller
            smaller = _MAGICINTS[smallidx - 1] // 2 if smallidx > _FIRSTIDX else 0
        elif is_smaller > 0:
            smaller = smallnum
            smallnum = _MAGICINTS[smallidx] // 2
        sizesmall = [_MAGICINTS[smallidx]] * 3
    return np.array(out[:natoms], dtype=np.int32)


def read_xtc(path):
    """Decode a GROMACS .xtc trajectory.

    Returns ``(xyz float32 (F,N,3), time float32 (F,), box float32 (F,3,3))``.
    Coordinates are ``float32(int) * float32(1/precision)`` -- the xdrfile
    convention; this exact rounding is what makes the mapping KAT bit-exact.
    """
    with open(path, "rb") as handle:
        blob = handle.read()
    pos = 0
    frames, times, boxes = [], [], []
    while pos < len(blob):
        magic, natoms, _step = struct.unpack_from(">iii", blob, pos)
        if magic != 1995:
            raise ValueError("not an XTC frame")
        (time,) = struct.unpack_from(">f", blob, pos + 12)
        box = np.frombuffer(blob, dtype=">f4", count=9, offset=pos + 16).astype(np.float32).reshape(3, 3)
        (natoms2,) = struct.unpack_from(">i", blob, pos + 52)
        pos += 56
        if natoms2 <= 9:
            xyz = np.frombuffer(blob, dtype=">f4", count=3 * natoms2, offset=pos).astype(np.float32)
            pos += 12 * natoms2
        else:
            (precision,) = struct.unpack_from(">f", blob, pos)
            minint = struct.unpack_from(">iii", blob, pos + 4)
            maxint = struct.unpack_from(">iii", blob, pos + 16)
            smallidx, nbytes = struct.unpack_from(">ii", blob, pos + 28)
            pos += 36
            ints = _decode_coords(natoms2, minint, maxint, smallidx, blob[pos:pos + nbytes])
            pos += (nbytes + 3) // 4 * 4
            inv = np.float32(1.0) / np.float32(precision)
            xyz = (ints.astype(np.float32) * inv).reshape(-1)
        frames.append(xyz.reshape(natoms, 3))
        times.append(time)
        boxes.append(box)
    return (np.array(frames, dtype=np.float32), np.array(times, dtype=np.float32),
            np.array(boxes, dtype=np.float32))

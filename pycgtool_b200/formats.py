"""GRO / XTC file formats.

The reference delegates file I/O to mdtraj (``pycgtool/frame.py:41-63,231-244``),
which is not available here; this module covers what the hot path's callers
need: GROMACS ``.gro`` (read / write, with mdtraj's water renaming) and ``.xtc``
(read / write).  The XTC codec lives in the shared library (``csrc/cgb_xtc.cu``):
``read_xtc`` decodes on the host through it, ``read_xtc_device`` uploads the
*compressed* file image and decodes on the GPU, one thread per frame, straight
into the device-resident trajectory the mapping kernel reads.
"""

import ctypes

import numpy as np

from . import _lib
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
    res_seq, serial = [], []
    xyz = np.empty((1, natoms, 3), dtype=np.float32)
    if natoms:
        dots = [i for i, ch in enumerate(body[0]) if ch == "." and i >= 20]
        width = dots[1] - dots[0] if len(dots) > 1 else 8
    for i, line in enumerate(body):
        key = (line[0:5], line[5:10])
        if key != last:
            resname = RESIDUE_NAME_REPLACEMENTS.get(line[5:10].strip(), line[5:10].strip())
            residues.append((resname, []))
            res_seq.append(int(line[0:5]))
            last = key
        serial.append(int(line[15:20]))
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
    top.res_seq, top.atom_serial = np.array(res_seq, dtype=np.int64), np.array(serial, dtype=np.int64)
    return Trajectory(xyz, top, np.zeros(1, dtype=np.float32), box if box.any() else None)


def read_pdb(path):
    """First model of a .pdb -> ``Trajectory`` (Angstrom -> nm, float32).  Fixed-column ATOM / HETATM
    records and CRYST1; residues change with (chain, resSeq, iCode, resName).  A CRYST1 with zero
    lengths means "no box", as mdtraj treats it."""
    residues, last, coords = [], None, []
    box = None
    with open(path) as handle:
        for line in handle:
            rec = line[:6]
            if rec == "CRYST1":
                a, b, c = (float(line[6 + 9 * k:15 + 9 * k]) / 10.0 for k in range(3))
                angles = [float(line[33 + 7 * k:40 + 7 * k]) for k in range(3)]
                if a > 0 and b > 0 and c > 0:
                    box = _box_from_lengths_angles(a, b, c, *angles)
            elif rec in ("ATOM  ", "HETATM"):
                key = (line[21], line[22:27], line[17:21].strip())
                if key != last:
                    resname = RESIDUE_NAME_REPLACEMENTS.get(key[2], key[2])
                    residues.append((resname, []))
                    last = key
                resname = residues[-1][0]
                atom = line[12:16].strip()
                residues[-1][1].append(ATOM_NAME_REPLACEMENTS.get(resname, {}).get(atom, atom))
                coords.append([float(line[30:38]) / 10.0, float(line[38:46]) / 10.0, float(line[46:54]) / 10.0])
            elif rec == "ENDMDL":
                break
    if not coords:
        raise OSError(f"'{path}': PDB format is not supported (no ATOM records)")
    xyz = np.asarray(coords, dtype=np.float32)[None]
    top = Topology.from_residue_lists(residues)
    return Trajectory(xyz, top, np.zeros(1, dtype=np.float32), None if box is None else box[None])


def _box_from_lengths_angles(a, b, c, alpha, beta, gamma):
    """Row vectors of the cell (mdtraj.utils.lengths_and_angles_to_box_vectors)."""
    alpha, beta, gamma = np.radians([alpha, beta, gamma])
    va = [a, 0.0, 0.0]
    vb = [b * np.cos(gamma), b * np.sin(gamma), 0.0]
    cx = c * np.cos(beta)
    cy = c * (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / np.sin(gamma)
    cz = np.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
    box = np.array([va, vb, [cx, cy, cz]], dtype=np.float64)
    box[np.abs(box) < 1e-6 * max(a, b, c)] = 0.0
    return box.astype(np.float32)


def write_gro(path, traj, frame=0):
    """Write one frame in mdtraj's .gro dialect (the file's own numbering for topologies read from a
    file, zero-based for topologies built in memory; 3 decimals, always nine box fields)."""
    top = traj.topology
    xyz = traj.xyz[frame]
    res_seq = top.res_seq if top.res_seq is not None else np.arange(top.n_residues)
    serial = top.atom_serial if top.atom_serial is not None else np.arange(top.n_atoms)
    lines = ["Generated with MDTraj, t= %s" % float(traj.time[frame]), " %d" % top.n_atoms]
    res_of_atom = np.repeat(np.arange(top.n_residues), np.diff(top.res_first))
    for i in range(top.n_atoms):
        r = int(res_of_atom[i])
        lines.append("%5d%-5s%5s%5d%8.3f%8.3f%8.3f" % (
            int(res_seq[r]) % 100000, top.res_names[top.res_name_id[r]], top.atom_names[top.atom_name_id[i]],
            int(serial[i]) % 100000, xyz[i, 0], xyz[i, 1], xyz[i, 2]))
    v = traj.unitcell_vectors
    b = np.zeros((3, 3)) if v is None else v[frame]
    lines.append(("%10.5f" * 9) % (b[0, 0], b[1, 1], b[2, 2], b[0, 1], b[0, 2], b[1, 0], b[1, 2], b[2, 0], b[2, 1]))
    with open(path, "w") as handle:
        handle.write("\n".join(lines) + "\n")


# ------------------------------------------------------------------------------- XTC
def _scan_xtc(image):
    """Frame table, boxes and times of an XTC file image (numpy uint8) through ``cgb_xtc_scan``."""
    lib = _lib.load()
    n_frames, n_atoms = ctypes.c_int64(), ctypes.c_int32()
    _lib.check(lib.cgb_xtc_scan(_lib.ptr(image), image.size, 0, None, None, None, None,
                                ctypes.byref(n_frames), ctypes.byref(n_atoms)), "cgb_xtc_scan")
    nf = n_frames.value
    table = np.zeros(nf * _lib.XTC_FRAME_BYTES, dtype=np.uint8)
    box = np.zeros((nf, 3, 3), dtype=np.float32)
    time = np.zeros(nf, dtype=np.float32)
    _lib.check(lib.cgb_xtc_scan(_lib.ptr(image), image.size, nf, _lib.ptr(table), _lib.ptr(box), _lib.ptr(time),
                                None, ctypes.byref(n_frames), ctypes.byref(n_atoms)), "cgb_xtc_scan")
    return table, box, time, nf, n_atoms.value


def _check_topology(topology, n_atoms):
    if topology is not None and topology.n_atoms != n_atoms:
        raise ValueError("Number of atoms does not match between topology and trajectory files")


def read_xtc(path, topology=None):
    """.xtc -> ``Trajectory`` decoded on the host (C codec in the shared library).
    Coordinates are ``float32(int) * float32(1/precision)``."""
    image = np.fromfile(path, dtype=np.uint8)
    try:
        table, box, time, nf, n_atoms = _scan_xtc(image)
    except RuntimeError as exc:
        raise OSError(f"'{path}': XTC format is not supported ({exc})") from exc
    _check_topology(topology, n_atoms)
    xyz = np.empty((nf, n_atoms, 3), dtype=np.float32)
    padded = np.concatenate([image, np.zeros(8, dtype=np.uint8)])
    rc = _lib.load().cgb_xtc_decode_host(_lib.ptr(padded), _lib.ptr(table), nf, n_atoms, _lib.ptr(xyz))
    if rc == -1:
        raise OSError(f"'{path}': truncated or corrupt XTC coordinate stream")
    _lib.check(rc, "cgb_xtc_decode_host")
    return Trajectory(xyz, topology, time, box if box.any() else None)


XTC_FRAME_DTYPE = np.dtype([("offset", "<i8"), ("nbytes", "<i4"), ("natoms", "<i4"), ("minint", "<i4", 3),
                            ("maxint", "<i4", 3), ("smallidx", "<i4"), ("precision", "<f4")])
assert XTC_FRAME_DTYPE.itemsize == _lib.XTC_FRAME_BYTES


def read_image_pinned(path):
    """An .xtc file read straight into pinned host memory and scanned: ``(pinned uint8 tensor of the image plus 8
    bytes of padding, size, frame table, box vectors (F, 3, 3), time (F,), n_frames, n_atoms)``."""
    import os
    import torch
    size = os.path.getsize(path)
    pinned = torch.empty(size + 8, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    with open(path, "rb") as handle:
        got = handle.readinto(memoryview(host)[:size])
    if got != size:
        raise OSError(f"'{path}': short read")
    host[size:] = 0
    try:
        table, box, time, nf, n_atoms = _scan_xtc(host[:size])
    except RuntimeError as exc:
        raise OSError(f"'{path}': XTC format is not supported ({exc})") from exc
    return pinned, size, table, box, time, nf, n_atoms


def read_xtc_device(path, topology=None, device=None):
    """.xtc -> ``Trajectory`` whose coordinates are decoded on the GPU and stay there.

    Only the compressed file image crosses PCIe (read straight into pinned memory, uploaded in blocks of
    frames while the previous block is being decoded); ``cgb_xtc_decode`` expands it in HBM."""
    import torch
    from . import device as dev
    device = device or dev.default_device()
    pinned, size, table, box, time, nf, n_atoms = read_image_pinned(path)
    _check_topology(topology, n_atoms)
    xyz = torch.empty((nf, n_atoms, 3), dtype=torch.float32, device=device)
    bad = torch.zeros(1, dtype=torch.int32, device=device)
    decode_image_device(pinned, size, table, nf, n_atoms, xyz, bad_frames=bad)
    torch.cuda.current_stream().synchronize()      # the staging buffers may be released now
    if int(bad.item()):
        raise OSError(f"'{path}': {int(bad.item())} frame(s) hold a truncated or corrupt coordinate stream")
    return Trajectory(None, topology, time, box if box.any() else None, xyz_device=xyz)


def decode_image_device(pinned, n_bytes, table, n_frames, n_atoms, xyz_dev, block_bytes=128 << 20, on_block=None,
                        image_dev=None, n_streams=4, bad_frames=None, device=None):
    """Upload an XTC file image from pinned host memory and decode it on the GPU, pipelined: the image goes
    up in blocks of whole frames (~``block_bytes`` each) on a copy stream, and each block is decoded on one of
    ``n_streams`` side streams as soon as it has arrived.  A block's decode starts with a sequential walk per
    frame (a few ms of latency whatever the number of frames), so several blocks are kept in flight to hide
    it; the upload is then the limit.  ``pinned``: uint8 tensor of the image plus 8 bytes of padding;
    ``table``: the frame table from ``cgb_xtc_scan`` (numpy uint8).  ``on_block(f0, f1)`` is called with the
    block's stream current, after the decode of frames [f0, f1) has been enqueued (e.g. to map them right
    away).  The caller's stream waits for everything on return (asynchronously).  Returns the device image.

    ``xyz_dev=None`` streams: the coordinates are never held as a whole -- every pipeline slot decodes its block
    into a buffer of its own and ``on_block(f0, f1, coords)`` must consume ``coords`` (e.g. map it) on the block's
    stream before the slot's next block overwrites it (stream order guarantees that).  A trajectory whose raw
    floats exceed the GPU's memory flows through this way; only the compressed image (about a third) is resident."""
    import torch
    from . import device as dev
    device = xyz_dev.device if xyz_dev is not None else (device or dev.default_device())
    if image_dev is None:
        image_dev = torch.empty(n_bytes + 8, dtype=torch.uint8, device=device)
    table_dev = dev.to_device(table, device)
    if n_frames == 0:
        return image_dev
    recs = table.view(XTC_FRAME_DTYPE)
    offsets = recs["offset"].astype(np.int64)
    # block boundaries: frame f1 starts a new block once the block holds >= block_bytes
    starts = [0]
    while starts[-1] < n_frames:
        target = offsets[starts[-1]] + block_bytes
        nxt = int(np.searchsorted(offsets, target, side="left"))
        nxt = min(max(nxt, starts[-1] + 1), starts[-1] + 65535, n_frames)
        starts.append(nxt)
    main = torch.cuda.current_stream(device)
    copy = _side_stream(device, "copy")
    workers = [_side_stream(device, i) for i in range(max(1, min(n_streams, len(starts) - 1)))]
    ring = None
    if xyz_dev is None:
        widest = max(f1 - f0 for f0, f1 in zip(starts[:-1], starts[1:]))
        ring = [torch.empty((widest, n_atoms, 3), dtype=torch.float32, device=device) for _ in workers]
    entry = torch.cuda.Event()
    entry.record(main)
    copy.wait_event(entry)             # buffers may still be in use by earlier work on the caller's stream
    b0 = 0
    for k, (f0, f1) in enumerate(zip(starts[:-1], starts[1:])):
        b1 = int(offsets[f1]) if f1 < n_frames else n_bytes + 8
        with torch.cuda.stream(copy):
            image_dev[b0:b1].copy_(pinned[b0:b1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        worker = workers[k % len(workers)]
        with torch.cuda.stream(worker):
            if k < len(workers):
                worker.wait_event(entry)
            worker.wait_event(ready)
            block_out = xyz_dev[f0:f1] if ring is None else ring[k % len(workers)][:f1 - f0]
            decode_device(image_dev, n_bytes, table_dev, f1 - f0, n_atoms, block_out, frame0=f0,
                          workspace_key=k % len(workers), bad_frames=bad_frames)
            if on_block is not None:
                if ring is None:
                    on_block(f0, f1)
                else:
                    on_block(f0, f1, block_out)
        b0 = b1
    for worker in workers:
        main.wait_stream(worker)
    held = [t for t in (image_dev, table_dev, xyz_dev, bad_frames) if t is not None] + (ring or [])
    for t in held:
        for st in workers + [copy]:
            t.record_stream(st)
    return image_dev


_SIDE_STREAMS = {}


def _side_stream(device, name):
    import torch
    key = (device.type, device.index, name)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def decode_device(image_dev, n_bytes, table_dev, n_frames, n_atoms, xyz_dev, block=4096, frame0=0, workspace_key=None,
                  bad_frames=None):
    """``cgb_xtc_decode`` over blocks of frames (one launch handles at most 65535 frames; the
    checkpoint workspace is sized for one block and reused).  Decodes frames [frame0, frame0 + n_frames)
    of the table into ``xyz_dev[0:n_frames]``.  ``bad_frames``: device int32 counter of frames whose bit stream
    is invalid (truncated or corrupt file); the caller reads it once the stream has finished."""
    from . import device as dev
    lib = _lib.load()
    block = int(min(block, max(n_frames, 1)))
    ws_bytes = int(lib.cgb_xtc_workspace_bytes(block, n_atoms))
    workspace = _workspace(xyz_dev.device, ws_bytes, workspace_key)
    frame_bytes = _lib.XTC_FRAME_BYTES
    for f0 in range(0, n_frames, block):
        f1 = min(n_frames, f0 + block)
        _lib.check(lib.cgb_xtc_decode(dev.dptr(image_dev), n_bytes, dev.dptr(table_dev[(frame0 + f0) * frame_bytes:]),
                                      f1 - f0, n_atoms, dev.dptr(xyz_dev[f0:]), dev.dptr(workspace), ws_bytes,
                                      dev.dptr(bad_frames), dev.stream_handle()), "cgb_xtc_decode")


_WORKSPACES = {}


def _workspace(device, n_bytes, slot=None):
    """Decoder scratch, cached per device and pipeline slot (decodes of one slot run in stream order)."""
    import torch
    key = (device.type, device.index, slot)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < n_bytes:
        ws = torch.empty(n_bytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def encode_xtc(xyz, unitcell_vectors=None, time=None, precision=1000.0):
    """(F, N, 3) float32 coordinates -> XTC file image (numpy uint8), via ``cgb_xtc_encode``."""
    lib = _lib.load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    nf, n_atoms = xyz.shape[0], xyz.shape[1]
    box = None if unitcell_vectors is None else np.ascontiguousarray(unitcell_vectors, dtype=np.float32).reshape(nf, 9)
    tm = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
    capacity = 4096 + nf * (128 + 8 * n_atoms)
    written = ctypes.c_int64()
    for _ in range(2):
        out = np.empty(capacity, dtype=np.uint8)
        rc = lib.cgb_xtc_encode(_lib.ptr(xyz), _lib.ptr(box), _lib.ptr(tm), nf, n_atoms, float(precision),
                                _lib.ptr(out), capacity, ctypes.byref(written))
        if rc == 0:
            return out[:written.value].copy()
        if rc != -2 or written.value <= capacity:
            _lib.check(rc, "cgb_xtc_encode")
        capacity = written.value
    raise RuntimeError("cgb_xtc_encode: buffer negotiation failed")


def write_xtc(path, traj, precision=1000.0):
    """Write every frame of ``traj`` as .xtc."""
    image = encode_xtc(traj.xyz, traj.unitcell_vectors, traj.time, precision)
    image.tofile(path)


def load(path, top=None, on_device=False):
    """Dispatch on extension, like ``mdtraj.load`` for the supported formats.  ``on_device`` decodes
    an .xtc on the GPU (``read_xtc_device``)."""
    suffix = str(path).lower().rsplit(".", 1)[-1]
    if suffix == "gro":
        return read_gro(path)
    if suffix == "pdb":
        return read_pdb(path)
    if suffix == "xtc":
        return read_xtc_device(path, top) if on_device else read_xtc(path, top)
    raise OSError(f"'{path}': format is not supported")

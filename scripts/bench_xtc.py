"""Time the XTC ingest path on the GPU box: compressed image -> H2D -> GPU decode -> mapping -> D2H beads.

usage: python scripts/bench_xtc.py [frames]
"""
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import Frame, Mapping, _lib, device, formats  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    sysm = synth.s2()
    dev = torch.device("cuda")
    lib = _lib.load()
    xyz = sysm.coordinates_device(frames, dev).cpu().numpy()
    box = sysm.box_vectors(frames)
    t0 = time.perf_counter()
    image = formats.encode_xtc(xyz, box, precision=1000.0)
    t_enc = time.perf_counter() - t0
    n_atoms = sysm.n_atoms
    print(f"encode (host, all cores): {frames * n_atoms / t_enc:.3e} atoms/s, {len(image) / frames / n_atoms:.2f} B/atom")
    table, bx, tm, nf, na = formats._scan_xtc(image)
    padded = np.concatenate([image, np.zeros(8, dtype=np.uint8)])
    out = np.empty((nf, na, 3), dtype=np.float32)
    t0 = time.perf_counter()
    _lib.check(lib.cgb_xtc_decode_host(_lib.ptr(padded), _lib.ptr(table), nf, na, _lib.ptr(out)))
    t_dec = time.perf_counter() - t0
    print(f"decode (host, all cores): {nf * na / t_dec:.3e} atoms/s")
    pinned = torch.from_numpy(padded).pin_memory()
    table_dev = device.to_device(table, dev)
    img_dev = torch.empty_like(pinned, device=dev)
    xyz_dev = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    stream = device.stream_handle()

    def decode_only():
        formats.decode_device(img_dev, image.size, table_dev, nf, na, xyz_dev)

    img_dev.copy_(pinned)
    for _ in range(2):
        decode_only()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        decode_only()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"decode (GPU kernel):   {nf * na / ms / 1e-3:.3e} atoms/s  ({ms:.2f} ms for {nf} frames)")
    assert np.array_equal(xyz_dev.cpu().numpy(), out)

    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, Options())
    plan = mapper.compile_plan(sysm.topology)
    lengths = np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1)
    cg_host = torch.empty((nf, plan.n_beads, 3), dtype=torch.float32, pin_memory=True)

    def pipeline():
        img_dev.copy_(pinned, non_blocking=True)
        decode_only()
        cg = mapper.map_device(plan, xyz_dev, lengths)
        cg_host.copy_(cg, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        pipeline()
    t0 = time.perf_counter()
    for _ in range(3):
        pipeline()
    dt = (time.perf_counter() - t0) / 3
    print(f"file image -> beads on host (H2D {len(image) / 1e6:.0f} MB + decode + map + D2H): "
          f"{nf * na / dt:.3e} atom-frames/s ({dt * 1e3:.1f} ms)")


if __name__ == "__main__":
    main()

"""Sweep the K1 tuning knobs on a synthetic config and print achieved GB/s (run on the GPU box).

usage: python scripts/sweep_k1.py [s2|s3|s4] [frames]
"""
import itertools
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import Mapping, _lib  # noqa: E402
from workloads import synth  # noqa: E402
import pycgtool_b200.mapping as pm  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "s2"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    sysm = {"s2": synth.s2, "s3": synth.small_molecule, "s4": synth.s4}[which]()
    dev = torch.device("cuda")
    lib = _lib.load()
    xyz = sysm.coordinates_device(frames, dev)
    box = sysm.box_vectors(frames)
    lengths = torch.from_numpy(np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1)).to(dev)
    map_path, _ = sysm.write_files()
    results = []
    grid = [
        # (tile beads, span atoms, stage KB, stages, frames/CTA, consumer threads, FG)
        (32, 512, 24, 4, 128, 96, 4), (32, 512, 24, 3, 128, 96, 4), (32, 512, 20, 4, 256, 96, 4),
        (32, 512, 40, 3, 128, 192, 4), (32, 512, 40, 4, 256, 192, 4),
        (64, 900, 40, 3, 128, 192, 4), (64, 900, 40, 4, 128, 192, 4), (64, 900, 48, 4, 256, 192, 4),
        (64, 900, 20, 4, 128, 192, 2), (64, 900, 24, 6, 128, 192, 2),
        (96, 1300, 56, 3, 128, 256, 4), (96, 1300, 32, 4, 128, 256, 2), (96, 1300, 64, 3, 256, 256, 4),
        (36, 512, 24, 4, 128, 128, 4), (512, 512, 24, 4, 128, 128, 4),
        (32, 512, 12, 6, 128, 96, 2), (32, 512, 10, 8, 128, 96, 2), (64, 900, 20, 8, 256, 192, 2),
    ]
    if len(sys.argv) > 3:
        grid = [tuple(int(x) for x in arg.split(",")) for arg in sys.argv[3:]]
    for beads, span, stage_kb, nst, fpc, threads, fg in grid:
        pm.MAX_TILE_BEADS, pm.MAX_SPAN_ATOMS = beads, span
        mapper = Mapping(map_path, Options())
        plan = mapper.compile_plan(sysm.topology)
        out = torch.empty((frames, plan.n_beads, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.cgb_map_set_tuning(stage_kb * 1024, nst, fpc))
        _lib.check(lib.cgb_map_set_shape(threads, fg))
        try:
            for _ in range(3):
                mapper.map_device(plan, xyz, lengths, out=out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            reps = 5
            for _ in range(reps):
                mapper.map_device(plan, xyz, lengths, out=out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            nbytes = frames * (12 * sysm.n_atoms + 12 * plan.n_beads + 12)
            d = Mapping._device_plan(plan, dev)
            row = dict(beads=beads, span=span, stage_kb=stage_kb, stages=nst, fpc=fpc, threads=threads, fg=fg,
                       tiles=d["n_tiles"], max_span=d["max_span"], ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1))
        except RuntimeError as exc:
            row = dict(beads=beads, span=span, stage_kb=stage_kb, stages=nst, fpc=fpc, threads=threads, fg=fg,
                       error=str(exc)[:80])
        results.append(row)
        print(json.dumps(row), flush=True)
    lib.cgb_map_set_tuning(0, 0, 0)
    lib.cgb_map_set_shape(0, 0)
    best = max((r for r in results if "gbs" in r), key=lambda r: r["gbs"])
    print("BEST", json.dumps(best))


if __name__ == "__main__":
    main()

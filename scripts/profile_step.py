#!/usr/bin/env python
"""Host-side profile of one north-star step (map_and_measure + boltzmann_invert) on a quarter-size S5 membrane.

    python scripts/profile_step.py [frames] [reps]
"""
import cProfile
import pathlib
import pstats
import sys
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda")
    sysm = synth.s5(scale=0.25)
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(frames, dev, chunk=8)
    box = sysm.box_vectors(frames)
    mapper, bonds = Mapping(map_path, Options()), BondSet(bnd_path, Options())
    frame = Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz)
    buf = torch.empty((frames, mapper._plan_for(frame._topology).n_beads, 3), dtype=torch.float32, device=dev)

    def step():
        map_and_measure(mapper, bonds, frame, store_values=False, out=buf)
        bonds.boltzmann_invert()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    st = bonds._state
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        st.packed.zero_()
        st.relaunch()
    stop.record()
    torch.cuda.synchronize()
    print(f"step {wall * 1e3:.3f} ms, kernels {start.elapsed_time(stop) / reps:.3f} ms")
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    prof.disable()
    pstats.Stats(prof).sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()

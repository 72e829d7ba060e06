#!/usr/bin/env python
"""Run the fused mapping + measurement pass a few times on a quarter-size S5 membrane (profiling target).

    python scripts/run_pipeline_once.py [frames] [passes]
"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda")
    sysm = synth.s5(scale=0.25)
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(frames, dev, chunk=8)
    box = sysm.box_vectors(frames)
    mapper, bonds = Mapping(map_path, Options()), BondSet(bnd_path, Options())
    frame = Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz)
    map_and_measure(mapper, bonds, frame, fused=True)
    st = bonds._state
    for _ in range(passes):
        st.packed.zero_()
        st.relaunch()
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()

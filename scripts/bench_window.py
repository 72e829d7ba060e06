"""Streamed mapping of a solvated membrane whose solvent is not mapped (run on the GPU box):
host-resident trajectory -> beads on the host, with and without the atom window."""
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import pycgtool_b200.mapping as pm  # noqa: E402
from pycgtool_b200 import Frame, Mapping  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    sysm = synth.membrane(576, 100000, seed=5)            # S2's lipids + 1.2 M solvent atoms
    frames = 256
    dev = torch.device("cuda")
    xyz = torch.empty((frames, sysm.n_atoms, 3), dtype=torch.float32, pin_memory=True)
    xyz.copy_(sysm.coordinates_device(frames, dev, chunk=16))
    box = sysm.box_vectors(frames)
    map_path, _ = sysm.write_files()
    map_path.write_text(map_path.read_text().replace("[ W4 ]", "[ NOTW4 ]"))
    mapper = Mapping(map_path, Options())
    pm.STREAM_THRESHOLD_BYTES = 0
    frame = Frame.from_arrays(sysm.topology, xyz.numpy(), unitcell_vectors=box)
    plan = mapper._plan_for(sysm.topology)
    first, held = Mapping.atom_window(plan)
    results = {}
    for label, window in (("window", (first, held)), ("whole frames", (0, sysm.n_atoms))):
        plan._atom_window = window
        for _ in range(2):
            out = mapper.apply(frame)._trajectory.xyz
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = mapper.apply(frame)._trajectory.xyz
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        results[label] = out
        print(f"{label:13s}: {held if label == 'window' else sysm.n_atoms:9d} of {sysm.n_atoms} atoms uploaded per frame, "
              f"{1e3 * dt:7.2f} ms for {frames} frames = {frames * sysm.n_atoms / dt:.3e} atom-frames/s")
    assert np.array_equal(results["window"], results["whole frames"])


if __name__ == "__main__":
    main()

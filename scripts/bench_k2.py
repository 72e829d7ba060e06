"""Time the measurement kernels (K2) alone on a synthetic config (run on the GPU box).

usage: python scripts/bench_k2.py [s3|s5small] [frames]
"""
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import BondSet, Frame, Mapping, _lib, device  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


STAGED = "--gather" not in sys.argv


def main():
    if "--gather" in sys.argv:
        sys.argv.remove("--gather")
    tune = [a for a in sys.argv if a.startswith("--tune=")]
    for a in tune:
        sys.argv.remove(a)
    which = sys.argv[1] if len(sys.argv) > 1 else "s3"
    dev = torch.device("cuda")
    lib = _lib.load()
    if tune:
        sb, ns = (int(x) for x in tune[0][7:].split(","))
        _lib.check(lib.cgb_measure_set_tuning(sb, ns), "cgb_measure_set_tuning")
        print(f"tuning: stage_bytes={sb} stages={ns}")
    if which == "s3":
        sysm, frames = synth.small_molecule(), int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    elif which == "s5":
        sysm, frames = synth.s5(), int(sys.argv[2]) if len(sys.argv) > 2 else 16
    else:
        sysm, frames = synth.s5(scale=0.25), int(sys.argv[2]) if len(sys.argv) > 2 else 64
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(frames, dev, chunk=1 << 14 if which == "s3" else 8)
    box = sysm.box_vectors(frames)
    cg = Mapping(map_path, Options()).apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    del xyz
    bonds = BondSet(bnd_path, Options())
    t0 = time.perf_counter()
    bonds.apply(cg)
    torch.cuda.synchronize()
    print(f"first apply (plan + upload + kernels): {1e3 * (time.perf_counter() - t0):.2f} ms")
    t0 = time.perf_counter()
    bonds.apply(cg)
    torch.cuda.synchronize()
    print(f"second apply (API level): {1e3 * (time.perf_counter() - t0):.2f} ms")
    st = bonds._state
    plan = st.plan
    n2, n3, n4 = plan.counts
    d = BondSet._device_plan(plan, dev)
    cgx = cg._trajectory.device_xyz()
    box_dev = device.to_device(box.reshape(frames, 9), dev)
    values = st.values
    stream = device.stream_handle()

    hist_lo, hist_hi = torch.from_numpy(st.hist_lo).to(dev), torch.from_numpy(st.hist_hi).to(dev)

    def run(c2, c3, c4, hist=True, vals=True):
        parts = torch.zeros_like(st.partials)
        h = torch.zeros_like(st.hist) if hist else None
        # column counts can only be switched off from the tail, so time cumulative prefixes
        _lib.check(lib.cgb_measure(device.dptr(cgx), device.dptr(box_dev), 1, frames, plan.n_beads,
                                   device.dptr(d["col_beads"]), device.dptr(d["col_bond"]), device.dptr(d["col_slot"]),
                                   plan.max_slots, device.dptr(d["tile_span"]) if STAGED else None, plan.max_span,
                                   c2, c3, c4, plan.n_bonds, device.dptr(st.shift), device.dptr(values) if vals else None,
                                   device.dptr(parts), device.dptr(h), device.dptr(hist_lo), device.dptr(hist_hi), 128, stream))

    def timed(label, *args, **kw):
        for _ in range(2):
            run(*args, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            run(*args, **kw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{label:40s} {ms:8.3f} ms")
        return ms

    # NOTE: with fewer columns than the plan's, the values row stride (M) shrinks too; timing only
    t_all = timed("lengths+angles+dihedrals, hist, values", n2, n3, n4)
    timed("same, no histogram", n2, n3, n4, hist=False)
    timed("same, no histogram, no values", n2, n3, n4, hist=False, vals=False)
    timed("lengths only", n2, 0, 0)
    timed("lengths+angles", n2, n3, 0)
    m = frames * (n2 + n3 + n4)
    touched = len(np.unique(plan.col_beads))
    nbytes = frames * (12 * touched + 4 * (n2 + n3 + n4) + 36)   # 12 B per bead touched
    print(f"measures {m:.3e}  {m / t_all / 1e-3:.3e} measures/s  algorithmic {nbytes / 1e9:.3f} GB -> "
          f"{nbytes / t_all / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()

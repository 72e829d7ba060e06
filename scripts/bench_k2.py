#!/usr/bin/env python
"""K2 micro-benchmark: the fused measurement kernel (cgb_measure_fused) on S3 and on a quarter-size S5 membrane.

    python scripts/bench_k2.py [s3|membrane|both] [--sweep] [--frames N]

For every mode (values + histograms / statistics only / no histogram) prints one JSON line: ms per pass, measures/s,
algorithmic GB/s (12 B per bead touched + 4 B per stored measure + 36 B of box per frame) and the fraction
of the measured HBM peak.  --sweep also walks the per-call tuning struct (stage bytes, stages, run length, waves).
"""

import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import BondSet, Frame, Mapping, _lib  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def peak():
    path = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(path.read_text())["hbm_gbs"]) if path.exists() else 6650.0


def setup(workload, frames, dihedrals=False):
    dev = torch.device("cuda")
    if workload == "s3":
        sysm, n_frames, chunk = synth.small_molecule(), frames or (1 << 20), 1 << 16
    else:
        sysm, n_frames, chunk = synth.s5(scale=0.25), frames or 256, 8
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(n_frames, dev, chunk=chunk)
    box = sysm.box_vectors(n_frames)
    opts = Options()
    opts.generate_dihedrals = dihedrals
    cg = Mapping(map_path, opts).apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    del xyz
    return cg, BondSet(bnd_path, opts), n_frames


def time_mode(cg, bonds, n_frames, store, hist, tuning=None, reps=5):
    bonds.tuning = tuning
    bonds.apply(cg, histogram=hist, store_values=store)
    st = bonds._state
    plan = st.plan
    nbins = int(bonds.histogram_bins) if hist else 0

    def launch():
        st.packed.zero_()
        bonds._launch_measure(plan, st.source, n_frames, st.shift, st.values, st.packed, nbins, st.hist_lo, st.hist_hi)

    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        launch()
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps
    n_terms = plan.n_cols
    touched = len(np.unique(plan.col_beads))
    nbytes = n_frames * (12 * touched + (4 * n_terms if store else 0) + 36)
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"ms": round(ms, 4), "measures_per_s": n_frames * n_terms / (ms * 1e-3), "gbs": round(gbs, 1),
            "frac": round(gbs / peak(), 4), "store": store, "hist": hist, "terms": n_terms, "frames": n_frames,
            "tiles": len(plan.tiles), "groups": len(plan.groups)}


def main():
    args = sys.argv[1:]
    which = args[0] if args and not args[0].startswith("-") else "both"
    sweep = "--sweep" in args
    frames = int(args[args.index("--frames") + 1]) if "--frames" in args else 0
    for workload in (["s3", "membrane"] if which == "both" else [which]):
        for dihedrals in ([False] if workload == "s3" else [False, True]):
            cg, bonds, n_frames = setup(workload, frames, dihedrals)
            for store, hist in ((True, True), (False, True), (True, False), (False, False)):
                row = time_mode(cg, bonds, n_frames, store, hist)
                row.update(workload=workload, dihedrals=dihedrals)
                print(json.dumps(row), flush=True)
            if sweep:
                for sb in (8192, 16384, 24576, 32768):
                    for ns in (2, 3, 4):
                        for run in (256, 1024):
                            for waves in (1, 2):
                                tuning = _lib.MeasureTuning(stage_bytes=sb, n_stages=ns, max_run=run, ctas_per_sm=waves)
                                try:
                                    row = time_mode(cg, bonds, n_frames, True, True, tuning, reps=3)
                                except RuntimeError as exc:
                                    row = {"error": str(exc)[:80]}
                                row.update(workload=workload, dihedrals=dihedrals, stage_bytes=sb, n_stages=ns,
                                           max_run=run, ctas_per_sm=waves)
                                print(json.dumps(row), flush=True)
            del cg, bonds
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

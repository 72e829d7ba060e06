#!/usr/bin/env python
"""Benchmark of the PyCGTOOL hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   the reference algorithm on the host CPUs

Workload (BASELINE.json configs[1], "S2"): map-only DOPC/POPC membrane, 1,152 lipids,
N = 156,672 atoms -> B = 13,824 MARTINI beads, 10,000 frames, synthetic coordinates
(molecules straddle the periodic box).  A *step* is one pass of the mapping over the whole
10,000-frame trajectory.  Metric: atom-frames/s.

* ``value``      device-resident throughput: inputs already in HBM, CUDA events around K steps.
* ``e2e``        the same metric through the public API (``Mapping.apply``) with HOST buffers:
                 pinned host -> device copies, the kernel and the device -> host copy of the bead
                 coordinates are all inside the timed region (on a bounded frame batch, stated).
* ``roofline``   HBM roofline of the dominant kernel (K1 ``map_tiles_kernel``): algorithmic bytes
                 F*(12N + 12B + 12) per launch / CUDA-event duration, against MEASURED_PEAKS.json.
* ``cpu_baseline`` the oracle (NumPy restatement of the reference, its loop structure kept) timed on
                 the host cores on a bounded frame sample.
* ``secondary``  the other half of BASELINE.json's metric, bond-measures/s: K2 (one ``measure_fused_kernel`` launch
                 for all kinds) on config S3 (1 M frames, 66 terms) with its own ``roofline``, ``cpu_baseline``
                 (oracle bonds.measure + invert) and ``e2e`` (``BondSet.apply`` + ``boltzmann_invert`` from a pinned
                 host CG trajectory); ``measure_membrane``: the same kernel on a quarter-size S5 membrane.
* ``pipeline``   SURVEY 8f-1: mapping + measurement as ONE kernel (``cgb_map_measure``) against K1 then K2.
* ``north_star_s5`` BASELINE.json configs[4]: S5 (20 M atoms) map + bond statistics, frames sharded over the
                 ranks, ONE NCCL all-reduce of the packed [moments | histograms] buffer and the Boltzmann epilogue
                 inside the timed step; strong (512 frames in total) and weak (256 frames per GPU) scaling.
* ``ingest``     compressed .xtc image -> GPU decode -> mapping -> beads on the host.

Multi-GPU (torchrun, one rank per GPU): frames are independent, so every rank maps its own
10,000-frame shard with no data-path collective (weak scaling); value = all atom-frames / max time.
"""

import argparse
import gc
import json
import math
import os
import pathlib
import statistics
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

S2_FRAMES = 10000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s2", choices=["s2", "s4", "s5"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0 = the config's own)")
    ap.add_argument("--e2e-frames", type=int, default=2048)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / e2e / measure legs")
    ap.add_argument("--ref-frames-per-worker", type=int, default=128)
    ap.add_argument("--legs", default="", help="comma list of supplementary legs: cpu,measure,membrane,pipeline,ingest,northstar")
    ap.add_argument("--northstar-scale", type=float, default=1.0, help="size of the S5 system of the north-star leg")
    return ap.parse_args()


def load_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = f"/tmp/cgb_clocks_{os.getpid()}.csv"

    def __enter__(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.out.close()

    def summary(self):
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                      parts[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
def make_system(name):
    from workloads import synth
    if name == "s2":
        return synth.s2(), S2_FRAMES, "S2 map-only DOPC/POPC membrane: 1,152 lipids, N=156,672 atoms -> 13,824 beads, 10,000 frames"
    if name == "s4":
        return synth.s4(), 2000, "S4 5M-atom solvated membrane: N=4,999,872 atoms -> 417,424 beads, 2,000 frames"
    return synth.s5(), 512, "S5 20M-atom multi-lipid membrane: N=19,999,992 atoms -> 1,710,442 beads, 512 frames"


def oracle_residue_lists(topology):
    top = topology
    return [(top.res_names[top.res_name_id[r]],
             [top.atom_names[i] for i in top.atom_name_id[top.res_first[r]:top.res_first[r + 1]]])
            for r in range(top.n_residues)]


def _oracle_map_chunk(args):
    """Worker of the reference arm: the oracle's Mapping.apply on one frame block."""
    map_path, residues, seed_sys, first, count = args
    from oracle import mapping as omap, system as osystem
    xyz = seed_sys.coordinates_host(count, first_frame=first)
    box = seed_sys.box_vectors(count, first_frame=first)
    sysm = osystem.from_residue_lists(residues, xyz, box)
    mols = omap.parse_map(map_path)
    t0 = time.perf_counter()
    cg = omap.apply(mols, sysm)
    dt = time.perf_counter() - t0
    return dt, float(cg.xyz[0, 0, 0])


def _oracle_bonds_chunk(args):
    """Worker of the reference arm, stage 2 + 3: oracle bonds.measure + invert on one block of S3 frames."""
    map_path, bnd_path, residues, seed_sys, first, count = args
    from oracle import bonds as obonds, mapping as omap, system as osystem
    xyz = seed_sys.coordinates_host(count, first_frame=first)
    box = seed_sys.box_vectors(count, first_frame=first)
    ocg = omap.apply(omap.parse_map(map_path), osystem.from_residue_lists(residues, xyz, box))
    terms = obonds.parse_bnd(bnd_path)
    t0 = time.perf_counter()
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    dt = time.perf_counter() - t0
    return dt, sum(len(t.indices) for mol in terms.values() for t in mol)


def reference_bonds_leg(pool, cores, frames_per_worker=4096):
    """bond-measures/s of the reference algorithm (oracle port) on all host cores: S3 frames split over workers."""
    from workloads import synth
    sysm = synth.small_molecule()
    map_path, bnd_path = sysm.write_files()
    residues = oracle_residue_lists(sysm.topology)
    best = None
    for rep in range(2):
        jobs = [(str(map_path), str(bnd_path), residues, sysm, (rep * cores + w) * frames_per_worker, frames_per_worker)
                for w in range(cores)]
        results = pool.map(_oracle_bonds_chunk, jobs)
        slowest = max(r[0] for r in results)
        best = slowest if best is None else min(best, slowest)
    n_terms = results[0][1]
    sample = (f"{cores * frames_per_worker} frames ({frames_per_worker} per worker x {cores} workers) of the "
              f"1,048,576-frame S3 config, oracle bonds.measure + invert")
    value = cores * frames_per_worker * n_terms / best
    return {"metric": "bond-measures/s", "value": value, "unit": "bond-measures/s",
            "cpu_baseline": {"value": value, "unit": "bond-measures/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "bond-measures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port -- mdtraj cannot be installed, so
    the reference itself cannot be imported) on the host cores, all of them, bounded frame sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    sysm, frames, label = make_system(args.workload)
    map_path, _ = sysm.write_files()
    residues = oracle_residue_lists(sysm.topology)
    cores = max(1, min(os.cpu_count() or 1, 64))
    per_worker = args.ref_frames_per_worker
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(str(map_path), residues, sysm, (step * cores + w) * per_worker % max(frames - per_worker, 1),
                     per_worker) for w in range(cores)]
            t0 = time.perf_counter()
            # coordinates are generated inside the workers before their timer starts; the wall time of
            # the step is the slowest worker's mapping time (workers run concurrently)
            results = pool.map(_oracle_map_chunk, jobs)
            wall = time.perf_counter() - t0
            slowest = max(r[0] for r in results)
            if step >= args.warmup:
                times.append(slowest)
            del wall
        secondary = None
        if not args.no_extras:
            try:
                secondary = reference_bonds_leg(pool, cores)
            except Exception as exc:  # noqa: BLE001 -- the headline line must still be printed
                secondary = {"error": repr(exc)}
    units = cores * per_worker * sysm.n_atoms
    total = sum(times)
    value = units * len(times) / total
    sample = f"{cores * per_worker} frames per step ({per_worker} per worker x {cores} workers) of the {frames}-frame config"
    line = {
        "metric": "atom-frames/s", "value": value, "unit": "atom-frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": label, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "atom-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if secondary is not None:
        line["secondary"] = secondary
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(sysm, map_path, frames_total, sample_frames=512):
    """Single-threaded oracle (the reference is single-threaded) on a bounded frame sample."""
    from oracle import mapping as omap, system as osystem
    residues = oracle_residue_lists(sysm.topology)
    xyz = sysm.coordinates_host(sample_frames)
    box = sysm.box_vectors(sample_frames)
    osys = osystem.from_residue_lists(residues, xyz, box)
    mols = omap.parse_map(map_path)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        omap.apply(mols, osys)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": sample_frames * sysm.n_atoms / best, "unit": "atom-frames/s", "cores": 1, "kind": "port",
            "sample": f"{sample_frames} of {frames_total} frames, single thread (the reference is single-threaded), "
                      f"best of 2, host has {os.cpu_count()} cores"}


def _bonds_cpu_baseline(sysm, map_path, bnd_path, frames_total, sample_frames):
    """Single-threaded oracle for stage 2 + 3: bonds.measure (the mdtraj kernels restated in NumPy) and the
    statistics / Boltzmann inversion of the reference (bondset.py:496-539, util.py:24-53) on a frame sample."""
    from oracle import bonds as obonds, mapping as omap, system as osystem
    residues = oracle_residue_lists(sysm.topology)
    xyz = sysm.coordinates_host(sample_frames)
    box = sysm.box_vectors(sample_frames)
    ocg = omap.apply(omap.parse_map(map_path), osystem.from_residue_lists(residues, xyz, box))
    terms = obonds.parse_bnd(bnd_path)
    t0 = time.perf_counter()
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    dt = time.perf_counter() - t0
    n_terms = sum(len(t.indices) for mol in terms.values() for t in mol)
    return {"value": sample_frames * n_terms / dt, "unit": "bond-measures/s", "cores": 1, "kind": "port",
            "sample": f"{sample_frames} of {frames_total} frames, oracle bonds.measure + invert, single thread, "
                      f"host has {os.cpu_count()} cores"}


def _k2_traffic(workload, n_frames):
    """DRAM bytes of one K2 launch from the committed ncu capture (per frame, scaled to this run), or None."""
    path = ROOT / "profiles" / "k2_traffic.json"
    try:
        tr = json.loads(path.read_text())[workload]
        return tr["dram_bytes_per_frame"] * n_frames, tr.get("source")
    except (OSError, ValueError, KeyError):
        return None, None


def measure_leg(dev, peak_gbs, workload="s3", with_cpu=True, with_e2e=True):
    """Stage 2 (K2, one cgb_measure_fused launch for all kinds): device resident, on config S3 or on a quarter-size
    S5 membrane.  `value` = values + moments + circular sums + 128-bin histograms in one pass; `stats_only` = the
    same without writing the values; `e2e` = BondSet.apply + boltzmann_invert from a pinned HOST CG trajectory."""
    import numpy as np
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, _lib, device
    from workloads import synth
    from workloads.synth import Options
    _lib.load()
    if workload == "s3":
        sysm, n_frames, chunk, cpu_frames = synth.small_molecule(), 1 << 20, 1 << 16, 1 << 15
        label = "S3 small molecule: 24 beads, 66 terms (23 bonds, 22 angles, 21 dihedrals), 1,048,576 frames"
    else:
        sysm, n_frames, chunk, cpu_frames = synth.s5(scale=0.25), 256, 8, 1
        label = "quarter-size S5 membrane (5M atoms), 256 frames"
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(n_frames, dev, chunk=chunk)
    box = sysm.box_vectors(n_frames)
    mapper = Mapping(map_path, Options())
    cg = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    del xyz
    bonds = BondSet(bnd_path, Options())
    for _ in range(2):
        bonds.apply(cg)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        bonds.apply(cg)
    torch.cuda.synchronize()
    api_ms = 1e3 * (time.perf_counter() - t0) / reps
    st = bonds._state
    plan = st.plan
    n2, n3, n4 = plan.counts
    n_terms = n2 + n3 + n4
    measures = n_frames * n_terms
    touched = len(np.unique(plan.col_beads))

    def timed(values, nbins):
        def launch():
            st.packed.zero_()
            bonds._launch_measure(plan, st.source, n_frames, st.shift, values, st.packed, nbins,
                                  st.hist_lo, st.hist_hi)
        for _ in range(2):
            launch()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            launch()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / reps

    ms = timed(st.values, 128)
    ms_stats = timed(None, 128)
    t0 = time.perf_counter()
    bonds.boltzmann_invert()
    torch.cuda.synchronize()
    invert_ms = 1e3 * (time.perf_counter() - t0)
    bytes_alg = n_frames * (12 * touched + 4 * n_terms + 36)   # 12 B per bead touched + 4 B per measure
    bytes_stats = n_frames * (12 * touched + 36)
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    traffic, traffic_src = _k2_traffic(workload, n_frames)
    if workload != "s3":
        label += f": {plan.n_beads:,} beads, {n_terms:,} terms per frame ({n2:,} bonds, {n3:,} angles, {n4:,} dihedrals)"
    leg = {"metric": "bond-measures/s", "value": measures / (ms * 1e-3), "unit": "bond-measures/s",
           "ms_per_step": ms, "dtype": "f32", "config": {"workload": label},
           "includes": "values + moments + circular sums + 128-bin histograms, ONE measure_fused_kernel launch "
                       "for all kinds",
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                        "traffic": traffic, "traffic_source": traffic_src, "kernel": "cgb::measure_fused_kernel",
                        "algorithmic_bytes_per_launch": bytes_alg, "frac_of_nominal_8TBs": gbs / 8000.0,
                        "limiter": "instruction issue (minimum image, atan2, statistics and histogram per "
                                   "measure), not DRAM"},
           "stats_only": {"value": measures / (ms_stats * 1e-3), "ms_per_step": ms_stats,
                          "algorithmic_bytes_per_launch": bytes_stats,
                          "achieved_gbs": bytes_stats / (ms_stats * 1e-3) / 1e9},
           "api_ms_per_pass": api_ms, "boltzmann_invert_ms": invert_ms, "gpu_launches_per_step": 1}
    if with_e2e:
        # e2e: the CG trajectory starts in pinned host memory; H2D, K2, the epilogue and the D2H of the fitted
        # parameters are inside the timed region (BondSet.apply + boltzmann_invert as a user calls them)
        ef = n_frames if workload == "s3" else min(n_frames, 128)
        host = torch.empty((ef, plan.n_beads, 3), dtype=torch.float32, pin_memory=True)
        host.copy_(cg._trajectory.device_xyz()[:ef])
        torch.cuda.synchronize()
        box_host = torch.from_numpy(np.ascontiguousarray(box[:ef], dtype=np.float32)).pin_memory()   # inputs are pinned
        hframe = Frame.from_arrays(cg._topology, host.numpy(), unitcell_vectors=box_host.numpy())

        def e2e_step():
            hframe._trajectory.drop_device()
            bonds.apply(hframe, store_values=False)
            bonds.boltzmann_invert()
            return bonds

        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        leg["e2e"] = {"value": ef * n_terms / dt, "unit": "bond-measures/s", "ms_per_step": dt * 1e3,
                      "h2d_bytes_per_step": int(host.numel() * 4 + ef * 36),
                      "d2h_bytes_per_step": int(plan.n_bonds * (8 * 5 + 8 * 128)), "frames_per_step": ef,
                      "note": "BondSet.apply(store_values=False) + boltzmann_invert on a pinned host CG trajectory: "
                              "H2D -> K2 -> epilogue -> D2H of parameters and histograms; PCIe-bound"}
        del host, hframe
    if with_cpu:
        leg["cpu_baseline"] = _bonds_cpu_baseline(sysm, map_path, bnd_path, n_frames, cpu_frames)
    return leg


def pipeline_leg(dev, peak_gbs, frames=64, workload="membrane"):
    """mapping + measurement in ONE kernel (cgb_map_measure) on a quarter-size S5 membrane, against the
    two separate launches (K1 then K2) on the same inputs.  Bytes: 12 N + 12 B + 4 M + 48 per frame."""
    import numpy as np
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure, pipeline
    from workloads import synth
    from workloads.synth import Options
    sysm = synth.s5(scale=0.25) if workload == "membrane" else synth.small_molecule()
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(frames, dev, chunk=8 if workload == "membrane" else 1 << 16)
    box = sysm.box_vectors(frames)
    mapper, bonds, bonds2 = Mapping(map_path, Options()), BondSet(bnd_path, Options()), BondSet(bnd_path, Options())
    frame = Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz)
    cg = map_and_measure(mapper, bonds, frame, fused=True)
    map_plan = mapper._plan_for(frame._topology)
    fused = pipeline._fused_plan(mapper, bonds, map_plan) is not None
    st = bonds._state
    plan = st.plan
    n_terms, n_atoms, n_beads = plan.n_cols, sysm.n_atoms, plan.n_beads
    reps = 5

    def run(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / reps

    def fused_step():
        st.packed.zero_()
        st.relaunch()

    ms_fused = run(fused_step) if fused else None
    lengths = torch.from_numpy(np.ascontiguousarray(np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1))).to(dev)
    out = cg._trajectory.device_xyz()
    bonds2.apply(cg)
    st2 = bonds2._state

    def separate_step():
        mapper.map_device(map_plan, xyz, lengths, out=out)
        st2.packed.zero_()
        bonds2._launch_measure(st2.plan, st2.source, frames, st2.shift, st2.values, st2.packed, 128,
                               st2.hist_lo, st2.hist_hi)

    ms_sep = run(separate_step)
    bytes_alg = frames * (12 * n_atoms + 12 * n_beads + 4 * n_terms + 48)
    ms = ms_fused if fused else ms_sep
    return {"what": "Mapping.apply + BondSet.apply as one pass over the atoms (map_and_measure): beads written once, "
                    "never re-read; values + statistics + histograms in the same kernel",
            "workload": f"{'quarter-size S5 membrane' if workload == 'membrane' else 'S3 small molecule'}: {n_atoms:,} atoms -> "
                        f"{n_beads:,} beads, {n_terms:,} terms, {frames} frames",
            "fused_kernel": fused, "ms_per_step": ms, "separate_k1_k2_ms": ms_sep,
            "atom_frames_per_s": frames * n_atoms / (ms * 1e-3), "bond_measures_per_s": frames * n_terms / (ms * 1e-3),
            "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": bytes_alg / (ms * 1e-3) / 1e9 / peak_gbs,
            "separate_achieved_gbs": (bytes_alg + frames * 12 * len(np.unique(plan.col_beads))) / (ms_sep * 1e-3) / 1e9}


def north_star_leg(dev, peak_gbs, world, rank, td, mode, scale=1.0):
    """BASELINE.json configs[4]: S5 (20 M atoms), full map + bond statistics, frames sharded over the ranks, the
    per-Bond moments and histograms combined by ONE NCCL all-reduce of the packed buffer (cgb_allreduce_partials).
    A step = map_and_measure (K1 + K2) -> all-reduce -> Boltzmann epilogue -> fitted parameters on the host:
    everything, the collective included, inside the timed region.  mode 'strong': 512 frames in total;
    'weak': 256 frames per GPU."""
    import numpy as np
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, dist, map_and_measure
    from workloads import synth
    from workloads.synth import Options
    sysm = synth.s5(scale=scale)
    total = 512 if mode == "strong" else 256 * world
    f0, f1 = dist.frame_shard(total, rank, world)
    frames = f1 - f0
    need = frames * (sysm.n_atoms * 12 + 2_000_000 * 12 * scale + 800_000 * 4 * scale) + (4 << 30)
    free = torch.cuda.mem_get_info(dev)[0]
    ok = torch.tensor([1 if need < free else 0], device=dev)
    if world > 1:
        td.all_reduce(ok, op=td.ReduceOp.MIN)
    if not int(ok.item()):
        return {"skipped": f"{mode}: {frames} frames per rank need {need / 1e9:.0f} GB, {free / 1e9:.0f} GB free"}
    map_path, bnd_path = sysm.write_files()
    xyz = sysm.coordinates_device(frames, dev, first_frame=f0, chunk=4)
    box = sysm.box_vectors(frames, first_frame=f0)
    mapper, bonds = Mapping(map_path, Options()), BondSet(bnd_path, Options())
    frame = Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz)

    cg_buffer = torch.empty((frames, mapper._plan_for(frame._topology).n_beads, 3), dtype=torch.float32, device=dev)

    def step():
        map_and_measure(mapper, bonds, frame, store_values=False, out=cg_buffer)   # block after block into one buffer
        bonds.boltzmann_invert()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    reps = 5
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        step()
    stop.record()
    torch.cuda.synchronize()
    t = torch.tensor([start.elapsed_time(stop) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item())
    # the kernel alone on this rank, and the collective alone (warm)
    st = bonds._state
    start.record()
    for _ in range(reps):
        st.packed.zero_()
        st.relaunch()
    stop.record()
    torch.cuda.synchronize()
    kernel_ms = start.elapsed_time(stop) / reps
    scratch = torch.zeros_like(st.packed)
    for _ in range(5):
        dist.allreduce_packed(scratch)
    torch.cuda.synchronize()
    start.record()
    for _ in range(20):
        dist.allreduce_packed(scratch)
    stop.record()
    torch.cuda.synchronize()
    coll_us = 1e3 * start.elapsed_time(stop) / 20
    plan = st.plan
    n_terms, n_atoms = plan.n_cols, sysm.n_atoms
    touched = len(np.unique(plan.col_beads))
    bytes_alg = frames * (12 * n_atoms + 12 * plan.n_beads + 12 * touched + 48)   # K1 + the beads K2 reads back
    return {"workload": f"S5 x{scale:g}: {n_atoms:,} atoms -> {plan.n_beads:,} beads, {n_terms:,} terms per frame, "
                        f"{plan.n_bonds} Bonds; {total} frames over {world} GPU(s) ({frames} on this rank)",
            "scaling": mode, "step": "map_and_measure (K1 then K2 on the resident beads, stats-only) -> NCCL all-reduce of "
                                     "[moments | histograms] -> Boltzmann epilogue -> parameters on the host",
            "ms_per_step": ms, "atom_frames_per_s": total * n_atoms / (ms * 1e-3),
            "bond_measures_per_s": total * n_terms / (ms * 1e-3), "kernel_ms": kernel_ms,
            "kernel_achieved_gbs": bytes_alg / (kernel_ms * 1e-3) / 1e9,
            "kernel_frac_of_hbm_peak": bytes_alg / (kernel_ms * 1e-3) / 1e9 / peak_gbs,
            "allreduce_warm_us": coll_us, "allreduce_bytes": int(st.packed.numel() * 8),
            "collective_inside_timed_region": True}


def ingest_leg(sysm, mapper, plan, dev, frames=2048, td=None, rank=0, world=1):
    """Supplementary: compressed .xtc image in pinned host memory -> H2D -> GPU decode
    (cgb_xtc_decode) -> K1 -> D2H of the bead coordinates, all inside the timed region.  With several ranks every
    rank pulls its own copy of the stream through its own GPU at the same time (rank 0 encodes the image once)."""
    import numpy as np
    import torch
    from pycgtool_b200 import _lib, device, formats
    lib = _lib.load()
    box = sysm.box_vectors(frames)
    shared = pathlib.Path("/dev/shm") / f"cgb_bench_xtc_{os.environ.get('MASTER_PORT', '0')}.img"
    if rank == 0:
        xyz = sysm.coordinates_device(frames, dev).cpu().numpy()
        image = formats.encode_xtc(xyz, box, precision=1000.0)
        del xyz
        if world > 1:
            image.tofile(shared)
    if world > 1:
        td.barrier()
        if rank != 0:
            image = np.fromfile(shared, dtype=np.uint8)
        td.barrier()
        if rank == 0:
            shared.unlink()
    n_atoms = sysm.n_atoms

    def run(image, box, host_check):
        table, _, _, nf, na = formats._scan_xtc(image)
        padded = np.concatenate([image, np.zeros(8, dtype=np.uint8)])
        host_rate = None
        sample = min(nf, 64)                  # host decoder timed on a bounded sample
        out = np.empty((sample, na, 3), dtype=np.float32)
        if host_check:
            t0 = time.perf_counter()
            _lib.check(lib.cgb_xtc_decode_host(_lib.ptr(padded), _lib.ptr(table), sample, na, _lib.ptr(out)))
            host_rate = sample * na / (time.perf_counter() - t0)
        pinned = torch.from_numpy(padded).pin_memory()
        table_dev = device.to_device(table, dev)
        img_dev = torch.empty(padded.shape, dtype=torch.uint8, device=dev)
        xyz_dev = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
        lengths = device.to_device(np.ascontiguousarray(np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1),
                                                        dtype=np.float32), dev)
        cg_host = torch.empty((nf, plan.n_beads, 3), dtype=torch.float32, pin_memory=True)
        d2h = torch.cuda.Stream(device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)

        def map_block(f0, f1):
            # beads of the block as soon as it is decoded; their D2H overlaps the next block's upload / decode
            cg = mapper.map_device(plan, xyz_dev[f0:f1], lengths[f0:f1])
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                cg_host[f0:f1].copy_(cg, non_blocking=True)
                cg.record_stream(d2h)

        def pipeline():
            formats.decode_image_device(pinned, image.size, table, nf, na, xyz_dev, on_block=map_block,
                                        image_dev=img_dev, bad_frames=bad)
            torch.cuda.synchronize()

        for _ in range(2):
            pipeline()
        assert int(bad.item()) == 0, "the GPU decoder rejected frames of a valid image"
        if host_check:
            assert np.array_equal(xyz_dev[:sample].cpu().numpy(), out), "GPU and host XTC decoders disagree"
        if world > 1:
            td.barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            pipeline()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:               # all ranks run at once: the slowest one sets the rate
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            dt = float(t.item())
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        formats.decode_device(img_dev, image.size, table_dev, nf, na, xyz_dev)
        stop.record()
        torch.cuda.synchronize()
        return nf, dt, nf * n_atoms / (start.elapsed_time(stop) * 1e-3), host_rate

    nf1, dt1, dec1, host_rate = run(image, box, True)
    torch.cuda.empty_cache()
    # the same image four times over: a stream long enough that the pipeline's fill and drain (upload of the first
    # block, walk + decode + map + D2H of the last) no longer weigh on the rate
    nf4, dt4, dec4, _ = run(np.tile(image, 4), np.tile(box, (4, 1, 1)), False)
    return {"what": "compressed .xtc image (pinned host) -> H2D -> GPU decode -> mapping -> D2H beads, pipelined "
                    "in 128 MB blocks of frames on 4 streams",
            "value": world * nf4 * n_atoms / dt4, "unit": "atom-frames/s", "frames": nf4, "ms": dt4 * 1e3,
            "ranks": world,
            "short_run": {"value": world * nf1 * n_atoms / dt1, "frames": nf1, "ms": dt1 * 1e3,
                          "note": f"fill + drain of the pipeline (one block's upload, walk, decode, map and D2H) cost "
                                  f"{1e3 * (dt1 - dt4 * nf1 / nf4):.1f} ms of this run"},
            "compressed_bytes_per_atom": len(image) / nf1 / n_atoms, "h2d_bytes": int(len(image)) * 4,
            "h2d_gbs_per_rank": 4 * len(image) / dt4 / 1e9,
            "gpu_decode_atoms_per_s": dec4, "host_decode_atoms_per_s": host_rate, "host_decode_cores": os.cpu_count(),
            "note": "decode = per-frame walk over a shared-memory ring (one warp per frame, all frames of a block "
                    "concurrently; branch-free steps, one LDS.64 per group) + parallel decode of checkpoints; upload of "
                    "block i+1 overlaps decode of block i"}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as td

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pycgtool_b200 import device as cgb_device
    cpus = cgb_device.bind_host_to_gpu(local) if world > 1 else None     # NUMA-local pinned buffers per rank
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner there when the communicator
        # is created, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=dev)
            td.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from pycgtool_b200 import Frame, Mapping, _lib
    from workloads.synth import Options
    lib = _lib.load()            # fails loudly if the CUDA library is missing

    peak_gbs, peak_src = load_peaks()
    sysm, frames, label = make_system(args.workload)
    if args.frames:
        frames = args.frames
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, Options())
    plan = mapper.compile_plan(sysm.topology)
    n_atoms, n_beads = sysm.n_atoms, plan.n_beads

    # each rank owns its own frame shard (different frames: first_frame offset), resident in HBM
    xyz = sysm.coordinates_device(frames, dev, first_frame=rank * frames)
    box_np = sysm.box_vectors(frames, first_frame=rank * frames)
    lengths = torch.from_numpy(np.stack([box_np[:, 0, 0], box_np[:, 1, 1], box_np[:, 2, 2]], axis=1)).to(dev)
    out = torch.empty((frames, n_beads, 3), dtype=torch.float32, device=dev)
    d = Mapping._device_plan(plan, dev)

    def step():
        mapper.map_device(plan, xyz, lengths, out=out)

    launches_per_step = (1 if d["n_tiles"] else 0) + (1 if d["n_wide"] else 0) + (1 if d["n_virt"] else 0)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # nvidia-smi samples every 100 ms and the K timed steps last tens of ms: run the same step
        # under the sampler before and after the timed region so that the load is actually observed
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.6:
            step()
            torch.cuda.synchronize()
        start.record()
        for _ in range(args.steps):
            step()
        stop.record()
        torch.cuda.synchronize()
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.6:
            step()
            torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    elapsed_ms = start.elapsed_time(stop)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    units_per_step = world * frames * n_atoms
    value = units_per_step / (ms_per_step * 1e-3)

    # roofline of the dominant kernel on this rank (one K1 launch per step)
    bytes_per_launch = frames * (12 * n_atoms + 12 * n_beads + 12)
    own_ms = start.elapsed_time(stop) / args.steps
    achieved = bytes_per_launch / (own_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": None, "kernel": "cgb::map_tiles_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    # DRAM traffic of one K1 launch from the committed ncu capture (dram__bytes_read.sum + write.sum),
    # recorded per frame of this workload so that it scales to the frame count of this run
    traffic_file = ROOT / "profiles" / "k1_traffic.json"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
            if tr.get("workload") == args.workload:
                roofline["traffic"] = tr["dram_bytes_per_frame"] * frames
                roofline["traffic_source"] = tr.get("source")
        except (ValueError, KeyError):
            pass

    # ---------------- e2e: public API with host buffers, copies inside the timed region
    e2e = None
    if not args.no_extras:
        ef = min(frames, args.e2e_frames)
        host = torch.empty((ef, n_atoms, 3), dtype=torch.float32, pin_memory=True)
        host.copy_(xyz[:ef])
        torch.cuda.synchronize()
        import pycgtool_b200.mapping as pm
        pm.STREAM_THRESHOLD_BYTES = 0
        frame = Frame.from_arrays(sysm.topology, host.numpy(), unitcell_vectors=box_np[:ef])
        for _ in range(2):
            cg = mapper.apply(frame)
            _ = cg._trajectory.xyz
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        reps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            cg = mapper.apply(frame)
            res = cg._trajectory.xyz            # host copy of the bead coordinates
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        own_gbs = (ef * n_atoms * 12 + ef * 12) / dt / 1e9
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        rates = [own_gbs]
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
            every = torch.zeros(world, dtype=torch.float64, device=dev)
            every[rank] = own_gbs
            td.all_reduce(every, op=td.ReduceOp.SUM)
            rates = [float(v) for v in every.tolist()]
        dt = float(t.item())
        e2e = {"value": world * ef * n_atoms / dt, "unit": "atom-frames/s", "h2d_gbs_per_rank": rates,
               "h2d_bytes_per_step": ef * n_atoms * 12 + ef * 12, "d2h_bytes_per_step": int(res.nbytes),
               "frames_per_step": ef, "ms_per_step": dt * 1e3,
               "note": "Mapping.apply on a pinned host trajectory: chunked H2D -> K1 -> D2H pipeline; PCIe-bound"}
        del host, frame, cg

    clock_info = clocks.summary()
    clock_info["window"] = "0.6 s of identical steps + the timed steps + 0.6 s of identical steps, 100 ms sampling"
    line = {
        "metric": "atom-frames/s", "value": value, "unit": "atom-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "b200",
        "config": {"workload": label, "frames_per_gpu": frames, "n_atoms": n_atoms, "n_beads": n_beads,
                   "l2": "inputs larger than L2 (%.1f GB per pass vs 126 MB), no flush needed" % (bytes_per_launch / 1e9)
                         if bytes_per_launch > 4 * 126e6 else
                         "inputs of %.0f MB per pass: comparable to the 126 MB L2 (reduced --frames run)" % (bytes_per_launch / 1e6),
                   "parallelism": f"frame-sharded x{world}, no data-path collective",
                   "host_binding": (f"rank bound to the {len(cpus)} CPU cores local to its GPU (NVML)" if cpus else
                                    "none")},
        "roofline": roofline, "clocks": clock_info, "gpu_launches": launches_per_step * args.steps,
    }
    if e2e is not None:
        line["e2e"] = e2e
    legs = set(args.legs.split(",")) if args.legs else {"cpu", "measure", "membrane", "pipeline", "ingest", "northstar"}
    if not args.no_extras:
        del xyz, out
        torch.cuda.empty_cache()

    def guarded(name, fn):
        """A supplementary leg must not take the headline down; with several ranks every rank still enters it."""
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001
            if world > 1:
                raise
            return {"error": f"{name}: {exc!r}"}

    if rank == 0 and world == 1 and not args.no_extras:
        if "cpu" in legs:
            line["cpu_baseline"] = cpu_baseline_leg(sysm, map_path, frames)
        if "measure" in legs:
            line["secondary"] = guarded("measure", lambda: measure_leg(dev, peak_gbs))
        if "membrane" in legs:
            torch.cuda.empty_cache()
            line["measure_membrane"] = guarded("membrane", lambda: measure_leg(dev, peak_gbs, "membrane", with_cpu=False))
        if "pipeline" in legs:
            torch.cuda.empty_cache()
            line["pipeline"] = guarded("pipeline", lambda: pipeline_leg(dev, peak_gbs))
            line["pipeline_s3"] = guarded("pipeline_s3", lambda: pipeline_leg(dev, peak_gbs, 1 << 18, "s3"))
        if "ingest" in legs and args.workload == "s2":
            torch.cuda.empty_cache()
            line["ingest"] = guarded("ingest", lambda: ingest_leg(sysm, mapper, plan, dev))
    if world > 1 and not args.no_extras and "ingest" in legs and args.workload == "s2":
        # the compressed path at N GPUs: 2.7x fewer bytes cross PCIe per atom than in the raw-float e2e above
        torch.cuda.empty_cache()
        line["ingest"] = ingest_leg(sysm, mapper, plan, dev, td=td, rank=rank, world=world)
    if world > 1 and not args.no_extras and "measure" in legs:
        # bond-measures/s at N GPUs (weak scaling): every rank measures its own 1,048,576-frame S3 shard (kernel
        # only; the collective is measured inside the north-star leg below)
        leg = measure_leg(dev, peak_gbs, with_cpu=False, with_e2e=False)
        t = torch.tensor([leg["ms_per_step"]], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        per_rank = leg["value"] * leg["ms_per_step"] * 1e-3
        leg["ms_per_step"] = float(t.item())
        leg["value"] = world * per_rank / (leg["ms_per_step"] * 1e-3)
        leg["scaling"] = "weak: one 1,048,576-frame shard per rank, max kernel time over ranks, no collective"
        leg.pop("roofline", None)
        leg.pop("stats_only", None)
        line["secondary"] = leg
    if not args.no_extras and "northstar" in legs:
        # BASELINE.json configs[4] at this N: every rank takes part (collective inside the timed region)
        torch.cuda.empty_cache()
        ns = {}
        for mode in ("strong", "weak"):
            gc.collect()                     # BondSet <-> state reference cycles hold the previous leg's trajectory
            torch.cuda.empty_cache()
            ns[mode] = guarded("northstar " + mode,
                               lambda m=mode: north_star_leg(dev, peak_gbs, world, rank, td, m, args.northstar_scale))
        line["north_star_s5"] = ns
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

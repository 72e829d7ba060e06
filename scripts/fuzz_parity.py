"""Randomised differential test of the CUDA path against the oracle (run on the GPU box).

usage: python scripts/fuzz_parity.py [seconds] [seed]

Every case draws a synthetic system (membrane patch with or without water, or a chain molecule), a frame
count, a box (large, tight -- many images -- or absent), a mapping centre, kernel tuning knobs and a frame
slice with an odd byte alignment, then requires: bead coordinates bit-identical to the oracle; lengths
bit-identical; angles / dihedrals within the float32 conditioning bound; statistics and fitted parameters
within 1e-4; histograms equal to numpy.histogram of the produced values.
"""
import math
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import bonds as obonds, geometry as ogeom, mapping as omap  # noqa: E402
from pycgtool_b200 import BondSet, Frame, Mapping, _lib  # noqa: E402
from workloads import synth  # noqa: E402
from pycgtool_b200.bondset import HIST_BINS, HIST_RANGES  # noqa: E402
from workloads.synth import Options  # noqa: E402
from tests.helpers import oracle_system  # noqa: E402

REL, ANGLE_ABS, COS_ULPS = 1e-4, 2e-6, 4e-7


def check_case(rng, case):
    lib = _lib.load()
    kind = rng.choice(["membrane", "membrane_water", "chain"])
    if kind == "chain":
        sysm = synth.small_molecule(n_beads=int(rng.integers(4, 40)), atoms_per_bead=int(rng.integers(1, 7)),
                                    seed=int(rng.integers(1 << 30)))
    else:
        sysm = synth.membrane(int(rng.integers(1, 60)), int(rng.integers(0, 400)) if kind == "membrane_water" else 0,
                              seed=int(rng.integers(1 << 30)))
    n_frames = int(rng.choice([1, 2, 3, 5, 8, 17, 33, 64, 129, 257, 513]))      # (chains fold their frames from ~8 folds on)
    box_mode = rng.choice(["large", "tight", "none", "triclinic"])
    if box_mode in ("tight", "triclinic"):
        sysm.box = np.array(rng.uniform(2.2, 3.5, 3))
    options = Options()
    # synthetic atom names (A0, A1, ...) carry no element, so masses cannot be guessed: geom / first only
    options.map_center = str(rng.choice(["geom", "first"]))
    options.generate_dihedrals = bool(rng.integers(2))
    import torch
    full = sysm.coordinates_host(n_frames + 1)
    xyz = np.ascontiguousarray(full[1:])
    # device-side slice: frame 0 of the call starts N * 12 bytes into the allocation (not 16-byte aligned in general)
    xyz_dev = torch.from_numpy(full).cuda()[1:]
    box = None if box_mode == "none" else sysm.box_vectors(n_frames, 1)
    if box_mode == "triclinic":          # mapping keeps using the lengths only; measurement searches 27 images
        box[:, 1, 0] = rng.uniform(-1.2, 1.2)
        box[:, 2, 0] = rng.uniform(-1.2, 1.2)
        box[:, 2, 1] = rng.uniform(-1.2, 1.2)
    map_path, bnd_path = sysm.write_files()
    # per-call tuning structs (the library keeps no state)
    tune = _lib.MapTuning(stage_bytes=int(rng.choice([0, 4096, 8192, 24576])), n_stages=int(rng.choice([0, 2, 3])),
                          frames_per_cta=int(rng.choice([0, 16, 128])))
    mtune = _lib.MeasureTuning(stage_bytes=int(rng.choice([0, 4096, 12288, 32768])), n_stages=int(rng.choice([0, 2, 3, 4])),
                               max_run=int(rng.choice([0, 8, 64])), ctas_per_sm=int(rng.choice([0, 1, 2])))
    store_values = bool(rng.integers(4))      # a quarter of the cases: statistics-only pass, values measured on demand
    if True:
        mapper = Mapping(map_path, options)
        mapper.tuning = tune
        cg = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz_dev))
        ocg = omap.apply(omap.parse_map(map_path, map_center=options.map_center),
                         oracle_system(sysm.topology, xyz, box))
        got = cg._trajectory.xyz
        assert got.shape == ocg.xyz.shape, "bead array shape"
        assert np.array_equal(got, ocg.xyz), "bead coordinates differ from the oracle"
        bondset = BondSet(bnd_path, options)
        bondset.tuning = mtune
        bondset.staged = bool(rng.integers(8))    # one case in eight through the gather kernels
        bondset.apply(cg, store_values=store_values)
        bondset.boltzmann_invert()
        terms = obonds.parse_bnd(bnd_path, generate_angles=options.generate_angles,
                                 generate_dihedrals=options.generate_dihedrals)
        obonds.measure(terms, ocg)
        obonds.invert(terms)
        hist = bondset._state.hist.cpu().numpy()
        k = 0
        n_terms = 0
        for mol in terms:
            for bond, term in zip(bondset[mol], terms[mol]):
                v, ref = np.asarray(bond.values), np.asarray(term.values)
                assert v.shape == ref.shape, "values shape"
                if len(term.atoms) == 2:
                    assert np.array_equal(v, ref), f"length {bond.atoms}"
                else:
                    diff = np.abs(v.astype(np.float64) - ref.astype(np.float64))
                    diff = np.minimum(diff, 2 * math.pi - diff)
                    tol = np.full(diff.shape, ANGLE_ABS)
                    if len(term.atoms) == 3:
                        sin = np.abs(np.sin(ref.astype(np.float64)))
                        tol += np.minimum(COS_ULPS / np.maximum(sin, 1e-12), math.sqrt(2 * COS_ULPS))
                    else:
                        q = term.indices
                        b1 = ogeom.displacement(ocg.xyz, q[:, 0], q[:, 1], ocg.box_vectors).astype(np.float64)
                        b2 = ogeom.displacement(ocg.xyz, q[:, 1], q[:, 2], ocg.box_vectors).astype(np.float64)
                        b3 = ogeom.displacement(ocg.xyz, q[:, 2], q[:, 3], ocg.box_vectors).astype(np.float64)
                        n = lambda a: np.linalg.norm(a, axis=-1)  # noqa: E731
                        cond = n(b1) * n(b2) ** 2 * n(b3) / np.maximum(n(np.cross(b2, b3)) * n(np.cross(b1, b2)), 1e-30)
                        tol += COS_ULPS * cond.reshape(-1)
                    if not (diff <= tol).all():
                        w = int(np.argmax(diff - tol))
                        q = term.indices[w % len(term.indices)]
                        f = w // len(term.indices)
                        x = ocg.xyz[f].astype(np.float64)
                        bx = None if ocg.box_vectors is None or box_mode == "triclinic" else \
                            np.diag(ocg.box_vectors[f]).astype(np.float64)
                        def mic(d):
                            return d if bx is None else d - bx * np.round(d / bx)
                        if len(q) == 3:
                            u, w2 = mic(x[q[0]] - x[q[1]]), mic(x[q[2]] - x[q[1]])
                            truth = math.acos(max(-1.0, min(1.0, float(u @ w2) / math.sqrt(float(u @ u) * float(w2 @ w2)))))
                        else:
                            truth = float("nan")
                        raise AssertionError(f"angle {bond.atoms} got {float(v[w])!r} oracle {float(ref[w])!r} float64 "
                                             f"{truth!r} tol {float(tol[w]):.3g}")
                lo, hi = HIST_RANGES[len(term.atoms)]
                want, _ = np.histogram(v[~np.isnan(v)].astype(np.float64), bins=HIST_BINS,
                                       range=(float(np.float32(lo)), float(np.float32(hi))))
                if not np.array_equal(hist[k], want):
                    bad = np.flatnonzero(hist[k] != want)
                    edges = np.linspace(float(np.float32(lo)), float(np.float32(hi)), HIST_BINS + 1)
                    detail = []
                    for b in bad[:4]:
                        near = v[(v > edges[max(b - 1, 0)]) & (v < edges[min(b + 2, HIST_BINS)])]
                        t = (near.astype(np.float64) - np.float64(np.float32(lo))) * HIST_BINS / (
                            np.float64(np.float32(hi)) - np.float64(np.float32(lo)))
                        frac = t - np.floor(t)
                        closest = near[np.argsort(np.minimum(frac, 1 - frac))[:3]]
                        detail.append((int(b), int(hist[k][b]), int(want[b]), [float(x) for x in closest],
                                       [float(x) for x in np.sort(np.minimum(frac, 1 - frac))[:3]]))
                    raise AssertionError(f"histogram {bond.atoms} n={len(v)} sum {hist[k].sum()} vs {want.sum()} "
                                         f"bins(bin, got, want, closest values, edge distance)={detail}")
                if term.eqm is not None and n_frames > 1:
                    # The statistics are checked against the oracle's statistics of the values the GPU produced:
                    # with a handful of frames and a small spread the variance is so ill-conditioned that the
                    # permitted last-ulp differences of the angles themselves move it by more than 1e-4.
                    from oracle import stats as ostats
                    want_eqm, want_fc = ostats.boltzmann_invert(v, len(term.atoms), term.form)
                    part = bondset._state.partials[k].cpu().numpy()
                    extra = ""
                    if len(v) <= 16:
                        extra = (f" values={[float(x) for x in v]} oracle_values={[float(x) for x in ref]} partials="
                                 f"{[float(x) for x in part]} sumcos={float(np.cos(v.astype(np.float64)).sum())!r} "
                                 f"sumsin={float(np.sin(v.astype(np.float64)).sum())!r}")
                    info = f"{bond.atoms} frames={n_frames} n={len(v)}{extra} got=({bond.eqm!r}, {bond.fconst!r}) " \
                           f"want=({want_eqm!r}, {want_fc!r}) oracle-on-oracle-values=({term.eqm!r}, {term.fconst!r})"
                    d_eqm = abs(bond.eqm - want_eqm)
                    if len(term.atoms) == 4:
                        d_eqm = min(d_eqm, 2 * math.pi - d_eqm)
                    slack = 0.0
                    if len(term.atoms) != 2:
                        # circular mean = direction of the resultant of unit vectors whose float32 components carry
                        # ~1e-7 of noise each: with nearly antipodal samples (resultant length R -> 0) the direction
                        # is ill-conditioned in the reference's own float32 arithmetic as well
                        vv = v[~np.isnan(v)].astype(np.float64)
                        resultant = math.hypot(np.cos(vv).sum(), np.sin(vv).sum()) / max(len(vv), 1)
                        slack = 4e-7 / max(resultant, 1e-9)
                    assert d_eqm <= REL * abs(want_eqm) + 2e-6 + slack, "eqm " + info
                    if math.isinf(want_fc):
                        assert bond.fconst == math.inf, "fconst " + info
                    else:
                        # the variance is taken about that mean: (d var) ~ 2 * mean-error * spread
                        assert abs(bond.fconst - want_fc) <= (REL + 50 * slack) * abs(want_fc), "fconst " + info
                k += 1
                n_terms += 1
        return kind, n_frames, box_mode, n_terms


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t0, case, failures = time.time(), 0, 0
    while time.time() - t0 < budget:
        state = rng.bit_generator.state
        try:
            info = check_case(rng, case)
            print(f"case {case:4d} ok   {info}", flush=True)
        except AssertionError as exc:
            failures += 1
            print(f"case {case:4d} FAIL {exc}", flush=True)
            if failures >= 5:
                break
        case += 1
    print(f"{case} cases, {failures} failures in {time.time() - t0:.0f} s")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()

"""GPU parity of stage 2 (measurement + statistics) and stage 3 (Boltzmann inversion).

Bar (north star): bond statistics and fitted parameters within 1e-4 relative of the
reference algorithm; index assignment bit-exact (checked on CPU in test_host_logic).
Lengths are additionally bit-exact against the oracle (same un-fused float32 ops).
"""

import json
import math

import numpy as np
import pytest

from oracle import bonds as obonds, geometry as ogeom, mapping as omap, stats as ostats, system as osystem
from tests.helpers import oracle_system

pytestmark = pytest.mark.gpu

REL = 1e-4            # north-star tolerance on statistics / parameters
ANGLE_ABS = 2e-6      # float32 acos / atan2 implementations differ in the last ulps (radians)
COS_ULPS = 4e-7       # the cosine fed to acos differs by a few float32 ulps (fused vs un-fused dot products)


def _acos_tolerance(theta):
    """Change of acos when its argument moves by COS_ULPS: d(cos) / sin(theta) away from the poles, and at most
    sqrt(2 d(cos)) next to them (cos = +-(1 - t^2 / 2) at distance t from a pole)."""
    sin = np.abs(np.sin(np.asarray(theta, dtype=np.float64)))
    return np.minimum(COS_ULPS / np.maximum(sin, 1e-12), math.sqrt(2 * COS_ULPS))


def _assert_values(bond, term, cg=None):
    """Lengths: bit-exact.  Angles / dihedrals: float32 acos / atan2 of quantities that differ by a
    few ulps (fused vs un-fused products), so the tolerance follows the conditioning of the term:
    d(theta) = d(cos) / sin(theta) for angles, and for dihedrals the rounding of the cross products
    relative to |c1||c2| (nearly collinear bead triples are ill-conditioned in any float32 kernel)."""
    got, ref = np.asarray(bond.values), np.asarray(term.values)
    assert got.shape == ref.shape and got.dtype == np.float32
    if len(term.atoms) == 2:
        assert np.array_equal(got, ref)
        return
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    diff = np.minimum(diff, 2 * math.pi - diff)
    tol = np.full(diff.shape, ANGLE_ABS)
    if len(term.atoms) == 3:
        tol = tol + _acos_tolerance(ref)
    elif cg is not None:
        q = term.indices
        b1 = ogeom.displacement(cg.xyz, q[:, 0], q[:, 1], cg.box_vectors).astype(np.float64)
        b2 = ogeom.displacement(cg.xyz, q[:, 1], q[:, 2], cg.box_vectors).astype(np.float64)
        b3 = ogeom.displacement(cg.xyz, q[:, 2], q[:, 3], cg.box_vectors).astype(np.float64)
        n = lambda v: np.linalg.norm(v, axis=-1)
        cond = n(b1) * n(b2) ** 2 * n(b3) / np.maximum(n(np.cross(b2, b3)) * n(np.cross(b1, b2)), 1e-30)
        tol = tol + COS_ULPS * cond.reshape(-1)
    else:
        tol = tol * 50
    assert (diff <= tol).all(), float((diff - tol).max())


def _assert_params(bond, term):
    if term.eqm is None:
        assert bond.eqm is None and bond.fconst is None
        return
    assert bond.eqm == pytest.approx(term.eqm, rel=REL, abs=2e-6)
    if math.isinf(term.fconst):
        assert bond.fconst == math.inf
    else:
        assert bond.fconst == pytest.approx(term.fconst, rel=REL)


def test_sugar_full_pipeline(golden, options):
    from pycgtool_b200 import BondSet, Frame, Mapping
    cg = Mapping(golden / "sugar.map", options).apply(Frame(golden / "sugar.gro", golden / "sugar.xtc"))
    measure = BondSet(golden / "sugar.bnd", options)
    measure.apply(cg)
    measure.boltzmann_invert()
    ocg = omap.apply(omap.parse_map(golden / "sugar.map"), osystem.load(golden / "sugar.gro", golden / "sugar.xtc"))
    terms = obonds.parse_bnd(golden / "sugar.bnd")
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    stored = json.loads((golden / "sugar_params.json").read_text())
    assert len(measure["ALLA"]) == 18
    for bond, term, row in zip(measure["ALLA"], terms["ALLA"], stored):
        assert len(bond.values) == 1001
        _assert_values(bond, term, ocg)
        _assert_params(bond, term)
        assert bond.eqm == pytest.approx(row["eqm"], rel=REL, abs=2e-6)
        assert bond.fconst == pytest.approx(row["fconst"], rel=REL)
        counts, edges = bond.histogram
        expect, _ = np.histogram(np.asarray(bond.values, dtype=np.float64), bins=len(counts),
                                 range=(float(np.float32(edges[0])), float(np.float32(edges[-1]))))
        assert np.array_equal(counts.astype(np.int64), expect)


def test_reference_measurement_kats(golden, options):
    # reference tests/test_bondset.py:64-85
    from pycgtool_b200 import BondSet, Frame
    measure = BondSet(golden / "sugar.bnd", options)
    measure.apply(Frame(golden / "sugar-cg.gro"))
    b = measure["ALLA"]
    assert len(b[0].values) == 1 and b[0].values[0] == pytest.approx(0.2225376, rel=1 / 500)
    assert len(b[6].values) == 1 and b[6].values[0] == pytest.approx(math.radians(77.22779289), rel=1 / 500)
    assert len(b[12].values) == 1 and b[12].values[0] == pytest.approx(math.radians(-89.5552903), rel=1 / 500)


@pytest.mark.parametrize("column,extra", [(1, {}), (2, {"default_fc": True}),
                                          (2, {"length_form": "MartiniDefaultLength",
                                               "angle_form": "MartiniDefaultAngle",
                                               "dihedral_form": "MartiniDefaultDihedral"})])
def test_reference_boltzmann_table(golden, options, column, extra):
    # reference tests/test_bondset.py:37-56,94-156, tolerance 0.5 %
    from pycgtool_b200 import BondSet, Frame, Mapping
    table = [(0.220474419132, 72658.18163, 1250), (0.221844516963, 64300.01188, 1250),
             (0.216313356504, 67934.93368, 1250), (0.253166204438, 19545.27388, 1250),
             (0.205958461836, 55359.06367, 1250), (0.180550961226, 139643.66601, 1250),
             (1.359314249473, 503.24211, 25), (2.026002746003, 837.76904, 25), (1.937848056214, 732.87969, 25),
             (1.453592079716, 945.32633, 25), (2.504189933347, 771.63691, 25), (1.733002945619, 799.82825, 25),
             (1.695540843207, 253.75691, 50), (-2.074156183261, 125.04591, 50), (2.768063749729, 135.50927, 50),
             (-2.213755550432, 51.13975, 50), (1.456495664734, 59.38162, 50), (-1.826134298998, 279.80889, 50)]
    for key, val in extra.items():
        setattr(options, key, val)
    measure = BondSet(golden / "sugar.bnd", options)
    cg = Mapping(golden / "sugar.map", options).apply(Frame(golden / "sugar.gro", golden / "sugar.xtc"))
    measure.apply(cg)
    measure.boltzmann_invert()
    for bond, row in zip(measure["ALLA"], table):
        assert bond.eqm == pytest.approx(row[0], rel=5e-3)
        assert bond.fconst == pytest.approx(row[column], rel=5e-3)


def test_polymer_neighbour_bonds_and_pbc(golden, options):
    # reference tests/test_bondset.py:158-182
    from pycgtool_b200 import BondSet, Frame
    bondset = BondSet(golden / "polyethene.bnd", options)
    bondset.apply(Frame(golden / "polyethene.gro"))
    assert [len(b.values) for b in bondset["ETH"]] == [5, 4, 4, 4]
    bondset.boltzmann_invert()
    assert bondset["ETH"][0].eqm == pytest.approx(0.107, rel=1 / 500)
    assert bondset["ETH"][1].eqm == pytest.approx(0.107, rel=1 / 500)
    terms = obonds.parse_bnd(golden / "polyethene.bnd")
    obonds.measure(terms, osystem.load(golden / "polyethene.gro"))
    obonds.invert(terms)
    for bond, term in zip(bondset["ETH"], terms["ETH"]):
        _assert_values(bond, term)
    bondset = BondSet(golden / "polyethene.bnd", options)
    bondset.apply(Frame(golden / "pbcpolyethene.gro"))
    bondset.boltzmann_invert()
    for bond in bondset.get_bond_lengths("ETH", True):
        assert bond.eqm == pytest.approx(1.0)
        assert bond.fconst == math.inf


def test_full_itp_sugar_and_vsites(golden, options, tmp_path):
    # reference tests/test_bondset.py:184-235 (sugar_out.itp predates circular statistics: rtol 5e-3)
    from pycgtool_b200 import BondSet, Frame, Mapping
    from pycgtool_b200.util import cmp_file_whitespace_float
    measure = BondSet(golden / "sugar.bnd", options)
    mapping = Mapping(golden / "sugar.map", options)
    measure.apply(mapping.apply(Frame(golden / "sugar.gro", golden / "sugar.xtc")))
    measure.boltzmann_invert()
    measure.write_itp(tmp_path / "sugar_out.itp", mapping)
    assert cmp_file_whitespace_float(tmp_path / "sugar_out.itp", golden / "sugar_out.itp", rtol=0.005, verbose=True)
    options.generate_angles = False
    measure = BondSet(golden / "martini3" / "naphthalene.bnd", options)
    mapping = Mapping(golden / "martini3" / "naphthalene.map", options)
    measure.apply(mapping.apply(Frame(golden / "martini3" / "naphthalene.gro")))
    measure.boltzmann_invert()
    measure.write_itp(tmp_path / "naphthalene_out.itp", mapping)
    assert cmp_file_whitespace_float(tmp_path / "naphthalene_out.itp",
                                     golden / "martini3" / "naphthalene_out.itp", rtol=0.005, verbose=True)


def test_dump_values_vs_alla_dat(golden, options, tmp_path):
    # reference tests/test_bondset.py:241-261
    from pycgtool_b200 import BondSet, Frame, Mapping
    from pycgtool_b200.util import cmp_file_whitespace_float
    measure = BondSet(golden / "sugar.bnd", options)
    measure.apply(Mapping(golden / "sugar.map", options).apply(Frame(golden / "sugar.gro", golden / "sugar.xtc")))
    measure.boltzmann_invert()
    measure.dump_values(output_dir=tmp_path)
    for name in ("ALLA_length.dat", "ALLA_angle.dat", "ALLA_dihedral.dat"):
        assert cmp_file_whitespace_float(golden / name, tmp_path / name, rtol=0.008, verbose=True)


def _synthetic_bonds(sysm, n_frames, options, ortho=True, skew=None):
    from pycgtool_b200 import BondSet, Frame, Mapping
    xyz = sysm.coordinates_host(n_frames)
    box = sysm.box_vectors(n_frames)
    if skew is not None:
        box[:, 1, 0] = skew[0]
        box[:, 2, 0] = skew[1]
        box[:, 2, 1] = skew[2]
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    bondset = BondSet(bnd_path, options)
    bondset.apply(cg)
    bondset.boltzmann_invert()
    ocg = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, box))
    terms = obonds.parse_bnd(bnd_path, generate_angles=options.generate_angles,
                             generate_dihedrals=options.generate_dihedrals)
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    return bondset, terms, ocg


def test_small_molecule_chain_statistics(options):
    """Config S3 shape (66 terms, n_inst = 1) on 4,096 frames, dihedrals straddling +-pi."""
    from workloads import synth
    bondset, terms, ocg = _synthetic_bonds(synth.small_molecule(), 4096, options)
    assert len(bondset["MOL"]) == 66
    for bond, term in zip(bondset["MOL"], terms["MOL"]):
        _assert_values(bond, term, ocg)
        _assert_params(bond, term)


def test_membrane_many_instances_ortho_pbc(options):
    from workloads import synth
    sysm = synth.membrane(40, 60, seed=17)
    sysm.box = np.array([4.0, 4.2, 3.9])
    sysm.sigma = 0.12
    options.generate_dihedrals = True
    bondset, terms, ocg = _synthetic_bonds(sysm, 37, options)
    for mol in ("DOPC", "POPC"):
        assert len(bondset[mol]) == 11 + 11 + len([t for t in terms[mol] if len(t.atoms) == 4])
        for bond, term in zip(bondset[mol], terms[mol]):
            assert len(bond.values) == 37 * 40
            _assert_values(bond, term, ocg)
            _assert_params(bond, term)


def test_triclinic_minimum_image(options):
    from workloads import synth
    sysm = synth.membrane(12, 0, seed=4)
    sysm.box = np.array([3.0, 3.0, 3.0])
    sysm.sigma = 0.1
    bondset, terms, ocg = _synthetic_bonds(sysm, 11, options, skew=(0.9, -1.1, 1.3))
    for mol in ("DOPC", "POPC"):
        for bond, term in zip(bondset[mol], terms[mol]):
            _assert_values(bond, term, ocg)
            _assert_params(bond, term)


def test_molecule_absent_and_custom_form(golden, options, tmp_path):
    """A molecule in the .bnd that is not in the frame keeps eqm/fconst None; a user-defined
    functional form receives the GPU statistics through the mean / variance hooks."""
    from pycgtool_b200 import BondSet, Frame, functionalforms

    class Quartic(functionalforms.FunctionalForm):
        gromacs_type_ids = (7, 7, 7)

        def fconst(self, values, temp):
            return 2.0 / self._variance_func(values) + self._mean_func(values)

    bnd = tmp_path / "x.bnd"
    bnd.write_text((golden / "sugar.bnd").read_text() + "\n[ GHOST ]\nA B\n")
    options.length_form = "Quartic"
    measure = BondSet(bnd, options)
    from pycgtool_b200 import Mapping
    cg = Mapping(golden / "sugar.map", options).apply(Frame(golden / "sugar.gro", golden / "sugar.xtc"))
    measure.apply(cg)
    measure.boltzmann_invert()
    ghost = measure["GHOST"][0]
    assert len(ghost.values) == 0 and ghost.eqm is None and ghost.fconst is None
    bond = measure["ALLA"][0]
    v = np.asarray(bond.values, dtype=np.float64)
    assert bond.gromacs_type_id == 7
    assert bond.fconst == pytest.approx(2.0 / v.var() + v.mean(), rel=REL)


def test_bond_level_invert_on_host_values(options):
    """``Bond.boltzmann_invert`` on caller-supplied values (reference tests/test_functionalforms.py flow)."""
    from pycgtool_b200.bondset import Bond
    from pycgtool_b200.functionalforms import CosHarmonic, Harmonic
    rng = np.random.default_rng(1)
    lengths = rng.normal(0.3, 0.01, size=5000).astype(np.float32)
    bond = Bond(("A", "B"), func_form=Harmonic())
    bond.values = lengths
    bond.boltzmann_invert(temp=300)
    eqm, fc = ostats.boltzmann_invert(lengths, 2, "Harmonic", 300)
    assert bond.eqm == pytest.approx(eqm, rel=REL) and bond.fconst == pytest.approx(fc, rel=REL)
    angles = rng.normal(2.0, 0.1, size=5000).astype(np.float32)
    bond = Bond(("A", "B", "C"), func_form=CosHarmonic())
    bond.values = list(angles)
    bond.boltzmann_invert()
    eqm, fc = ostats.boltzmann_invert(angles, 3, "CosHarmonic", 310)
    assert bond.eqm == pytest.approx(eqm, rel=REL) and bond.fconst == pytest.approx(fc, rel=REL)
    dih = (rng.normal(3.1, 0.3, size=5000) % (2 * np.pi) - np.pi + np.pi).astype(np.float32)
    dih = ((dih + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    bond = Bond(("A", "B", "C", "D"), func_form=Harmonic())
    bond.values = dih
    bond.boltzmann_invert()
    eqm, fc = ostats.boltzmann_invert(dih, 4, "Harmonic", 310)
    assert bond.eqm == pytest.approx(eqm, rel=REL, abs=2e-6) and bond.fconst == pytest.approx(fc, rel=REL)
    empty = Bond(("A", "B"), func_form=Harmonic())
    with pytest.raises(ValueError):
        empty.boltzmann_invert()


def test_frame_shards_merge_to_single_pass(options):
    """Two frame shards measured separately and summed (what the all-reduce does) equal one pass."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, dist
    from workloads import synth
    sysm = synth.small_molecule()
    n_frames = 600
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    mapper = Mapping(map_path, options)
    whole = BondSet(bnd_path, options)
    whole.apply(mapper.apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box)))
    shards = []
    for rank in range(2):
        b, e = dist.frame_shard(n_frames, rank, 2)
        part = BondSet(bnd_path, options)
        part.apply(mapper.apply(Frame.from_arrays(sysm.topology, xyz[b:e], unitcell_vectors=box[b:e])))
        shards.append(part)
    # same shift on every shard is what makes the sums addable; shard 1 re-measures with shard 0's
    assert torch.equal(shards[0]._state.shift, whole._state.shift)
    merged = dist.merge_partials([s._state.partials[:, :5] for s in shards]) if torch.equal(
        shards[1]._state.shift, whole._state.shift) else None
    counts = shards[0]._state.partials[:, 0] + shards[1]._state.partials[:, 0]
    assert torch.equal(counts, whole._state.partials[:, 0])
    hist = shards[0]._state.hist + shards[1]._state.hist
    assert torch.equal(hist, whole._state.hist)
    # circular sums do not depend on the shift and must add up to the single-pass sums
    circ = shards[0]._state.partials[:, 3:5] + shards[1]._state.partials[:, 3:5]
    assert torch.allclose(circ, whole._state.partials[:, 3:5], rtol=2e-6, atol=1e-3)
    assert merged is None or torch.allclose(merged, whole._state.partials[:, :5], rtol=2e-6, atol=1e-3)


def _chain_frame(n_beads, xyz, box_len):
    from pycgtool_b200 import Frame
    from pycgtool_b200.topology import Topology
    names = [f"B{i}" for i in range(n_beads)]
    top = Topology.from_residue_lists([("MOL", names)])
    n_frames = xyz.shape[0]
    box = np.zeros((n_frames, 3, 3), dtype=np.float32)
    box[:, 0, 0] = box[:, 1, 1] = box[:, 2, 2] = box_len
    lines = ["[ MOL ]"] + [f"{names[i]} {names[i + 1]}" for i in range(n_beads - 1)]
    lines += [f"{names[i]} {names[i + 1]} {names[i + 2]}" for i in range(n_beads - 2)]
    lines += [f"{names[i]} {names[i + 1]} {names[i + 2]} {names[i + 3]}" for i in range(n_beads - 3)]
    return top, Frame.from_arrays(top, xyz, unitcell_vectors=box), box, "\n".join(lines) + "\n"


@pytest.mark.parametrize("n_frames", [1, 2, 3, 13, 160])
def test_degenerate_and_nan_terms_follow_the_reference(options, tmp_path, n_frames):
    """Coincident beads (length 0, undefined angle), collinear triples, NaN coordinates and images across the
    box: values equal the oracle's (NaN where it is NaN), statistics skip NaN like np.nanmean / np.nanvar,
    histograms count what numpy.histogram counts.  Frame counts cover the pair loop, its tail and both."""
    from pycgtool_b200 import BondSet
    from pycgtool_b200.bondset import HIST_BINS, HIST_RANGES
    rng = np.random.default_rng(7 + n_frames)
    n_beads = 9
    xyz = (np.arange(n_beads)[None, :, None] * np.array([0.31, 0.07, -0.11]) +
           rng.normal(0, 0.12, (n_frames, n_beads, 3))).astype(np.float32)
    xyz[:, 5:] += np.float32(4.7)                    # the chain crosses the periodic boundary (box 5 nm)
    if n_frames > 2:
        xyz[2, 1] = xyz[2, 0]                        # coincident beads
    if n_frames > 5:
        xyz[5, 3] = np.nan                           # a NaN bead poisons every term it is part of
        xyz[7, 4:7] = xyz[7, 4] + np.outer(np.arange(3), np.array([0.25, 0.0, 0.0], dtype=np.float32))  # collinear
        xyz[9] = xyz[8]                              # a repeated frame
    top, frame, box, bnd_text = _chain_frame(n_beads, xyz, 5.0)
    bnd = tmp_path / "chain.bnd"
    bnd.write_text(bnd_text)
    bondset = BondSet(bnd, options)
    bondset.apply(frame)
    bondset.boltzmann_invert()
    terms = obonds.parse_bnd(bnd, generate_angles=False, generate_dihedrals=False)
    osys = oracle_system(top, xyz, box)
    obonds.measure(terms, osys)
    obonds.invert(terms)
    hist = bondset._state.hist.cpu().numpy()
    for k, (bond, term) in enumerate(zip(bondset["MOL"], terms["MOL"])):
        got, ref = np.asarray(bond.values), np.asarray(term.values)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (bond.atoms, got, ref)
        ok = ~np.isnan(ref)
        if len(term.atoms) == 2:
            assert np.array_equal(got[ok], ref[ok])
        else:
            # degenerate geometry (zero-length or collinear operands) is arbitrarily ill-conditioned: compare
            # the well-defined frames tightly and only require the degenerate ones to be finite where the oracle is
            diff = np.abs(got[ok].astype(np.float64) - ref[ok].astype(np.float64))
            diff = np.minimum(diff, 2 * math.pi - diff)
            regular = np.ones(ok.sum(), dtype=bool)
            frames_ok = np.flatnonzero(ok)
            for f in (2, 7):
                regular &= frames_ok != f
            assert (diff[regular] < 2e-4).all(), (bond.atoms, diff.max())
        lo, hi = HIST_RANGES[len(term.atoms)]
        # edges are float64 linspace of the float32 range limits (Python floats: numpy would otherwise round the
        # edges to float32)
        want, _ = np.histogram(got[~np.isnan(got)].astype(np.float64), bins=HIST_BINS,
                               range=(float(np.float32(lo)), float(np.float32(hi))))
        assert np.array_equal(hist[k], want), bond.atoms
        if n_frames >= 13 and not any(a in ("B0", "B1", "B4", "B5", "B6") for a in term.atoms):
            _assert_params(bond, term)


def test_s3_full_size_properties(options):
    """Config S3 at its full size (1,048,576 frames, 66 terms): sampled frames equal the oracle, and the
    size-independent properties hold -- every histogram holds exactly one count per in-range value, the
    streamed moments equal float64 moments of the stored values, and mapping then measuring a frame shard
    gives the values of the same frames in the full pass."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping
    from workloads import synth
    from pycgtool_b200.bondset import HIST_RANGES
    sysm = synth.small_molecule()
    n_frames = 1 << 20
    dev = torch.device("cuda")
    xyz = sysm.coordinates_device(n_frames, dev, chunk=1 << 14)
    box = sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    mapper = Mapping(map_path, options)
    cg = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    bondset = BondSet(bnd_path, options)
    bondset.apply(cg)
    bondset.boltzmann_invert()
    state = bondset._state
    hist = state.hist.cpu().numpy()
    partials = state.partials.cpu().numpy()
    shift = state.shift.cpu().numpy()
    bonds = bondset["MOL"]
    assert len(bonds) == 66
    for k in (0, 11, 22, 23, 40, 44, 45, 60, 65):          # lengths, angles and dihedrals
        bond = bonds[k]
        v = np.asarray(bond.values)
        assert v.shape == (n_frames,) and np.isfinite(v).all()
        lo, hi = HIST_RANGES[len(bond.atoms)]
        assert hist[k].sum() == np.count_nonzero((v >= np.float32(lo)) & (v <= np.float32(hi)))
        assert partials[k, 0] == n_frames
        d = v.astype(np.float64) - np.float64(shift[k])
        if len(bond.atoms) == 4:
            d -= 2 * math.pi * np.rint(d / (2 * math.pi))      # dihedrals: wrapped about the shift
        off = partials[k, 1] - d.sum()
        turns = np.rint(off / (2 * math.pi))       # a sample within an ulp of the antipode may wrap either way
        assert abs(turns) <= (3 if len(bond.atoms) == 4 else 0)
        assert abs(off - 2 * math.pi * turns) <= max(1e-5 * abs(d.sum()), 0.5)
        assert partials[k, 2] == pytest.approx((d * d).sum(), rel=1e-5)
        if len(bond.atoms) != 2:
            assert partials[k, 3] == pytest.approx(np.cos(v.astype(np.float64)).sum(), rel=1e-5, abs=0.5)
            assert partials[k, 4] == pytest.approx(np.sin(v.astype(np.float64)).sum(), rel=1e-5, abs=0.5)
    # sampled frames against the oracle
    pick = [0, 1, 77, 4095, 4096, 524287, n_frames - 2, n_frames - 1]
    sub = xyz[pick].cpu().numpy()
    ocg = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, sub, box[pick]))
    terms = obonds.parse_bnd(bnd_path, generate_angles=options.generate_angles,
                             generate_dihedrals=options.generate_dihedrals)
    obonds.measure(terms, ocg)
    for bond, term in zip(bonds, terms["MOL"]):
        class Sampled:
            values = np.asarray(bond.values)[pick]
        _assert_values(Sampled, term, ocg)


def test_s5_size_sampled_terms(options):
    """Largest configuration (S5, 1.7 M beads, 79 kinds of residue, 4 frames): for every Bond a handful of
    instances spread over the system is measured by the oracle on the same CG coordinates -- lengths bit-exact,
    angles within the float32 acos / atan2 conditioning -- and every value buffer has the reference's length."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping
    from workloads import synth
    sysm = synth.s5()
    n_frames = 4
    dev = torch.device("cuda")
    xyz = sysm.coordinates_device(n_frames, dev, chunk=1)
    box = sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    del xyz
    bondset = BondSet(bnd_path, options)
    bondset.apply(cg)
    plan = bondset._state.plan
    cg_xyz = cg._trajectory.xyz                                # (F, B, 3) on the host
    rng = np.random.default_rng(11)
    import itertools
    bonds = list(itertools.chain(*bondset._molecules.values()))
    assert len(bonds) == plan.n_bonds and plan.n_bonds >= 79
    checked = 0
    for k, bond in enumerate(bonds):
        table = plan.bond_indices[k]
        n_inst = len(table)
        if n_inst == 0:
            continue
        values = np.asarray(bond.values).reshape(n_frames, n_inst)
        pick = np.unique(np.concatenate([[0, n_inst - 1], rng.integers(0, n_inst, 6)]))
        ref = ogeom.MEASURE[len(bond.atoms)](cg_xyz, table[pick].astype(np.int64), box)
        got = values[:, pick]
        if len(bond.atoms) == 2:
            assert np.array_equal(got, ref), bond.atoms
        else:
            diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
            tol = ANGLE_ABS + _acos_tolerance(ref)
            assert (diff <= tol).all(), (bond.atoms, float(diff.max()))
        checked += got.size
    assert checked > 79 * 4 * 4


def test_measure_outputs_stay_inside_their_buffers(options):
    """compute-sanitizer is not available on this GPU pool, so the write bounds of K2 are checked with canaries:
    `values` and the packed `[partials | hist]` buffer live inside larger sentinel-filled buffers that must come
    back untouched (frame count that is not a multiple of 4, several tiles, frame blocks that do not divide the
    trajectory).  Called through the C ABI directly, with the plan the Python layer compiled."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, _lib, device
    from workloads import synth
    sysm = synth.membrane(70, 40, seed=23)
    n_frames = 37
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    bondset = BondSet(bnd_path, options)
    bondset.apply(cg)
    st = bondset._state
    plan = st.plan
    n_cols = plan.n_cols
    assert len(plan.tiles) >= 3 and plan.n_staged == n_cols
    dev = st.values.device
    d = BondSet._device_plan(plan, dev)
    guard = 4096
    n_packed = plan.n_bonds * (_lib.NPART + 128)
    vbuf = torch.full((n_frames * n_cols + 2 * guard,), -12345.0, dtype=torch.float32, device=dev)
    pbuf = torch.full((n_packed + 2 * guard,), -7.0, dtype=torch.float64, device=dev)
    values = vbuf[guard:guard + n_frames * n_cols]
    packed = pbuf[guard:guard + n_packed]
    packed.zero_()
    partials, hist = packed[:plan.n_bonds * _lib.NPART], packed[plan.n_bonds * _lib.NPART:]
    box_dev = cg._trajectory.device_box_vectors(dev)
    for tuning in (None, _lib.MeasureTuning(stage_bytes=8192, n_stages=2, max_run=8, ctas_per_sm=1)):
        values.fill_(-12345.0)
        packed.zero_()
        _lib.check(_lib.load().cgb_measure_fused(
            device.dptr(cg._trajectory.device_xyz()), device.dptr(box_dev), 1, n_frames, plan.n_beads,
            device.dptr(d["tiles"]), len(plan.tiles), device.dptr(d["groups"]), device.dptr(d["lanes"]),
            device.dptr(d["tile_bonds"]), device.dptr(d["tile_lists"]), plan.shape, plan.max_span, plan.max_slots,
            plan.max_edge_chunks, plan.max_tile_groups, 1, device.dptr(st.shift), device.dptr(values), n_cols,
            device.dptr(partials), device.dptr(hist),
            device.dptr(st.hist_lo), device.dptr(st.hist_hi), 128, _lib.tuning_ptr(tuning), device.stream_handle()))
        torch.cuda.synchronize()
        for buf, fill in ((vbuf, -12345.0), (pbuf, -7.0)):
            assert bool((buf[:guard] == fill).all()) and bool((buf[-guard:] == fill).all())
        assert torch.equal(values.view(n_frames, n_cols), st.values)    # and the same results as through the API
        assert torch.equal(hist.view(plan.n_bonds, 128), st.hist)
        assert torch.allclose(partials.view(plan.n_bonds, _lib.NPART), st.partials, rtol=1e-5, atol=1e-6)
        assert bool((values != -12345.0).all())


def test_gather_kernels_explicit_and_automatic(options, tmp_path):
    """The gather kernels (beads read from global memory) are the path for tiles whose bead range does not fit
    in shared memory.  (1) forced through `BondSet.staged = False` on a membrane; (2) reached automatically by a
    4,000-bead chain with terms that join its two ends (one frame of the tile's range is 48 KB)."""
    from pycgtool_b200 import BondSet, Frame, Mapping
    from workloads import synth
    sysm = synth.membrane(30, 20, seed=31)
    sysm.box = np.array([3.6, 3.3, 3.9])
    n_frames = 19
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    ocg = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, box))
    terms = obonds.parse_bnd(bnd_path, generate_angles=options.generate_angles,
                             generate_dihedrals=options.generate_dihedrals)
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    staged, gather = BondSet(bnd_path, options), BondSet(bnd_path, options)
    gather.staged = False
    for bondset in (staged, gather):
        bondset.apply(cg)
        bondset.boltzmann_invert()
    for mol in terms:
        for a, b, term in zip(staged[mol], gather[mol], terms[mol]):
            _assert_values(b, term, ocg)
            _assert_params(b, term)
            if len(term.atoms) == 2:
                assert np.array_equal(np.asarray(a.values), np.asarray(b.values))

    # (2) a chain whose first and last beads are bonded: the tile spans the whole 4,000-bead frame
    n_beads = 4000
    rng = np.random.default_rng(9)
    xyz = (np.cumsum(rng.normal(0, 0.2, (1, n_beads, 3)), axis=1) + rng.normal(0, 0.05, (7, n_beads, 3))).astype(np.float32)
    names = [f"B{i}" for i in range(n_beads)]
    from pycgtool_b200.topology import Topology
    top = Topology.from_residue_lists([("MOL", names)])
    box = np.zeros((7, 3, 3), dtype=np.float32)
    box[:, 0, 0] = box[:, 1, 1] = box[:, 2, 2] = 9.0
    lines = ["[ MOL ]"] + [f"B{i} B{i + 1}" for i in range(0, n_beads - 1, 7)] + ["B0 B3999", "B1 B3998"]
    lines += ["B0 B2000 B3999", "B5 B6 B7"] + ["B0 B1 B3998 B3999", "B10 B11 B12 B13"]
    bnd = tmp_path / "long.bnd"
    bnd.write_text("\n".join(lines) + "\n")
    frame = Frame.from_arrays(top, xyz, unitcell_vectors=box)
    bondset = BondSet(bnd, options)
    bondset.apply(frame)
    bondset.boltzmann_invert()
    assert bondset._state.plan.wide_counts == (2, 1, 1) and len(bondset._state.plan.tiles) > 0
    terms = obonds.parse_bnd(bnd, generate_angles=False, generate_dihedrals=False)
    osys = oracle_system(top, xyz, box)
    obonds.measure(terms, osys)
    obonds.invert(terms)
    for bond, term in zip(bondset["MOL"], terms["MOL"]):
        _assert_values(bond, term, osys)
        _assert_params(bond, term)


def test_statistics_are_reproducible_run_to_run(options):
    """No floating-point atomics inside a CTA (fixed-order float64 reduction); CTAs are combined with float64
    atomics, whose order can only move the last bits: repeated passes agree to 1e-13, histograms exactly."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping
    from workloads import synth
    sysm = synth.small_molecule()
    n_frames = 1 << 16
    xyz = sysm.coordinates_device(n_frames, torch.device("cuda"), chunk=1 << 14)
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(
        Frame.from_arrays(sysm.topology, unitcell_vectors=sysm.box_vectors(n_frames), xyz_device=xyz))
    runs = []
    for _ in range(3):
        bondset = BondSet(bnd_path, options)
        bondset.apply(cg)
        runs.append((bondset._state.partials.cpu().numpy().copy(), bondset._state.hist.cpu().numpy().copy()))
    for partials, hist in runs[1:]:
        assert np.array_equal(hist, runs[0][1])
        np.testing.assert_allclose(partials, runs[0][0], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("skew", [(0.9, -1.1, 1.3), (1.45, 1.45, 1.45), (2.2, -1.9, 2.6)])
def test_triclinic_lengths_equal_exhaustive_image_search(options, tmp_path, skew):
    """Independent check of the triclinic wrap (the reference has no fixture that wraps in a triclinic cell): the
    kernel's lengths against a float64 exhaustive search over periodic images (tests/bruteforce.py), which shares
    nothing with mdtraj's reduce-and-search algorithm that both the oracle and the kernel restate."""
    from pycgtool_b200 import BondSet, Frame
    from pycgtool_b200.topology import Topology
    from tests.bruteforce import min_image_distance, skewed_boxes
    rng = np.random.default_rng(21)
    n_frames, n_beads = 33, 40
    names = [f"B{i}" for i in range(n_beads)]
    top = Topology.from_residue_lists([("MOL", names)])
    box = skewed_boxes(n_frames, (3.0, 3.1, 2.9), skew, seed=5)
    frac = rng.uniform(-0.1, 1.1, size=(n_frames, n_beads, 3))
    xyz = np.einsum("fbi,fij->fbj", frac, box.astype(np.float64)).astype(np.float32)
    pairs = [(i, (5 * i + 3) % n_beads) for i in range(n_beads) if i != (5 * i + 3) % n_beads]
    pairs += [(i, i + 1) for i in range(0, n_beads - 1, 2)]
    pairs = sorted({tuple(sorted(p)) for p in pairs})
    bnd = tmp_path / "pairs.bnd"
    bnd.write_text("[ MOL ]\n" + "".join(f"B{a} B{b}\n" for a, b in pairs))
    options.generate_angles = False
    bonds = BondSet(bnd, options)
    bonds.apply(Frame.from_arrays(top, xyz, unitcell_vectors=box))
    want = min_image_distance(xyz, np.array(pairs), box, reach=4)
    assert len(bonds["MOL"]) == len(pairs)
    wrapped = 0
    for k, bond in enumerate(bonds["MOL"]):
        got = np.asarray(bond.values, dtype=np.float64)
        np.testing.assert_allclose(got, want[:, k], rtol=3e-6, atol=3e-6)
        direct = np.linalg.norm(xyz[:, pairs[k][1]].astype(np.float64) - xyz[:, pairs[k][0]], axis=-1)
        wrapped += int((want[:, k] < direct - 1e-3).sum())
    assert wrapped > 0.2 * n_frames * len(pairs)        # the image shift is exercised, not skipped


@pytest.mark.parametrize("n_frames,n_beads", [(4096 + 13, 24), (100, 24), (777, 9), (40, 24)])
def test_frame_folding_gives_the_unfolded_results(options, n_frames, n_beads):
    """Small systems measure several frames as one (frame folding): same values, statistics and histograms as the
    plain plan, including the unfolded tail (n_frames not a multiple of the fold), per-frame boxes and the dihedral
    second pass; a trajectory shorter than 8 folds is not folded at all."""
    from pycgtool_b200 import BondSet, Frame, Mapping
    from workloads import synth
    sysm = synth.small_molecule(n_beads=n_beads)
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    map_path, bnd_path = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    plain = BondSet(bnd_path, options)
    plain.fold_frames = False
    plain.apply(cg)
    plain.boltzmann_invert()
    folded = BondSet(bnd_path, options)
    folded.apply(cg)
    plan = folded._state.plan
    assert plan.fold > 1 and plan.folded is not None
    main = BondSet.folded_frames(plan, n_frames)
    assert main == (0 if n_frames < 8 * plan.fold else (n_frames // plan.fold) * plan.fold)
    folded.boltzmann_invert()
    stats_only = BondSet(bnd_path, options)
    stats_only.apply(cg, store_values=False)
    stats_only.boltzmann_invert()
    for a, b, c in zip(plain["MOL"], folded["MOL"], stats_only["MOL"]):
        assert np.array_equal(np.asarray(a.values), np.asarray(b.values), equal_nan=True)
        assert np.array_equal(np.asarray(a.values), np.asarray(c.values), equal_nan=True)   # re-measured on demand
        for other in (b, c):
            assert other.eqm == pytest.approx(a.eqm, rel=1e-6, abs=1e-7)
            assert other.fconst == pytest.approx(a.fconst, rel=2e-5)
            assert np.array_equal(other.histogram[0], a.histogram[0])

"""Pin the oracle against the reference's own golden files (CPU only).

Every expected number below comes from the reference repository: a golden file
under tests/data (copied to tests/golden/) or a known-answer value quoted from
one of its tests, cited next to the assertion.
"""

import json
import math
import re

import numpy as np
import pytest

from oracle import bonds, fileio, mapping, stats, system


@pytest.fixture(scope="module")
def sugar(golden):
    aa = system.load(golden / "sugar.gro", golden / "sugar.xtc")
    cg = mapping.apply(mapping.parse_map(golden / "sugar.map"), aa)
    terms = bonds.parse_bnd(golden / "sugar.bnd")
    bonds.measure(terms, cg)
    bonds.invert(terms)
    return aa, cg, terms["ALLA"]


def test_xtc_decoder_kat(golden):
    # reference tests/test_frame.py:53-56: first atom of water.xtc frame 0
    xyz, time, box = fileio.read_xtc(golden / "water.xtc")
    assert xyz.shape == (11, 663, 3)
    np.testing.assert_allclose(xyz[0, 0], [1.176, 1.152, 1.586], atol=1e-6)


def test_mapping_bit_exact_vs_sugar_out_xtc(sugar, golden):
    # reference tests/test_pycgtool.py:70-92 compares against this file with rtol 1e-3;
    # the restatement reproduces all 1001 x 6 x 3 floats exactly.
    _, cg, _ = sugar
    gold, _, _ = fileio.read_xtc(golden / "sugar_out.xtc")
    assert cg.xyz.shape == (1001, 6, 3)
    assert np.array_equal(cg.xyz, gold)


@pytest.mark.parametrize("kind,name,tol", [(2, "length", 6e-6), (3, "angle", 1e-4), (4, "dihedral", 2.5e-3)])
def test_measurements_vs_alla_dat(sugar, golden, kind, name, tol):
    # reference tests/test_bondset.py:241-261 (mdtraj outputs, 5 printed decimals; degrees for angular terms)
    _, _, terms = sugar
    ref = np.loadtxt(golden / f"ALLA_{name}.dat")
    mine = np.stack([t.values for t in terms if len(t.atoms) == kind], axis=1).astype(np.float64)
    if kind > 2:
        mine = np.degrees(mine)
    assert np.abs(mine - ref).max() < tol


def test_inversion_vs_out_rtp(sugar, golden):
    # reference tests/data/ffout.ff/out.rtp: current (circular) statistics, radians, 5 decimals
    _, _, terms = sugar
    rows = []
    for line in (golden / "ffout.ff" / "out.rtp").read_text().splitlines():
        toks = line.split()
        if len(toks) >= 4 and re.fullmatch(r"-?\d+\.\d{5}", toks[-1] if len(toks) < 7 else toks[-2]):
            nums = [float(t) for t in toks if re.fullmatch(r"-?\d+\.\d{5}", t)]
            rows.append(nums[:2])
    assert len(rows) == 18
    for term, (eqm, fc) in zip(terms, rows):
        assert abs(term.eqm - eqm) < 6e-6
        assert abs(term.fconst - fc) <= max(6e-6, 2e-7 * fc)


def test_sugar_params_fixture_matches_oracle(sugar, golden):
    _, _, terms = sugar
    stored = json.loads((golden / "sugar_params.json").read_text())
    for term, row in zip(terms, stored):
        assert list(term.atoms) == row["atoms"]
        assert term.eqm == pytest.approx(row["eqm"], rel=1e-12)
        assert term.fconst == pytest.approx(row["fconst"], rel=1e-12)


def test_bondset_apply_kats(golden):
    # reference tests/test_bondset.py:64-85
    cg = system.load(golden / "sugar-cg.gro")
    terms = bonds.parse_bnd(golden / "sugar.bnd")
    bonds.measure(terms, cg)
    t = terms["ALLA"]
    assert len(t) == 18 and all(len(x.values) == 1 for x in t)
    assert t[0].values[0] == pytest.approx(0.2225376, rel=1 / 500)
    assert math.degrees(t[6].values[0]) == pytest.approx(77.22779289, rel=1 / 500)
    assert math.degrees(t[12].values[0]) == pytest.approx(-89.5552903, rel=1 / 500)


def test_boltzmann_table(sugar):
    # reference tests/test_bondset.py:37-56,94-156 (0.5 %; table predates circular statistics)
    table = [(0.220474419132, 72658.18163), (0.221844516963, 64300.01188), (0.216313356504, 67934.93368),
             (0.253166204438, 19545.27388), (0.205958461836, 55359.06367), (0.180550961226, 139643.66601),
             (1.359314249473, 503.24211), (2.026002746003, 837.76904), (1.937848056214, 732.87969),
             (1.453592079716, 945.32633), (2.504189933347, 771.63691), (1.733002945619, 799.82825),
             (1.695540843207, 253.75691), (-2.074156183261, 125.04591), (2.768063749729, 135.50927),
             (-2.213755550432, 51.13975), (1.456495664734, 59.38162), (-1.826134298998, 279.80889)]
    for term, (eqm, fc) in zip(sugar[2], table):
        assert term.eqm == pytest.approx(eqm, rel=5e-3)
        assert term.fconst == pytest.approx(fc, rel=5e-3)


def test_polymer_instances_and_pbc(golden):
    # reference tests/test_bondset.py:158-182
    terms = bonds.parse_bnd(golden / "polyethene.bnd")
    bonds.measure(terms, system.load(golden / "polyethene.gro"))
    assert [len(t.values) for t in terms["ETH"]] == [5, 4, 4, 4]
    bonds.invert(terms)
    assert terms["ETH"][0].eqm == pytest.approx(0.107, rel=1 / 500)
    assert terms["ETH"][1].eqm == pytest.approx(0.107, rel=1 / 500)
    terms = bonds.parse_bnd(golden / "polyethene.bnd")
    bonds.measure(terms, system.load(golden / "pbcpolyethene.gro"))
    bonds.invert(terms)
    for t in terms["ETH"][:2]:
        assert t.eqm == pytest.approx(1.0, abs=1e-6)
        assert t.fconst == math.inf


def test_mapping_kats(golden):
    # reference tests/test_mapping.py:103-234
    water = system.load(golden / "water.gro")
    cg = mapping.apply(mapping.parse_map(golden / "water.map"), water)
    assert cg.natoms == 221
    np.testing.assert_allclose(cg.xyz[0, 0], [0.716, 1.300, 1.198], atol=5e-4)
    cg = mapping.apply(mapping.parse_map(golden / "water_double_weight.map"), water)
    np.testing.assert_allclose(cg.xyz[0, 0], [0.720, 1.294, 1.195], atol=5e-4)
    pbc = system.load(golden / "pbcwater.gro")
    cg = mapping.apply(mapping.parse_map(golden / "water.map"), pbc)
    np.testing.assert_allclose(cg.xyz[0, 0], pbc.xyz[0, 0])
    two = system.load(golden / "two.gro")
    for kwargs, expect, rtol in [({}, 1.5, 1e-6),
                                 ({"map_center": "mass", "itp_path": golden / "two.itp"}, 2.0, 1e-6),
                                 ({"map_center": "first", "itp_path": golden / "two.itp"}, 1.0, 1e-6),
                                 ({"map_center": "mass"}, 1.922575, 1e-2)]:
        cg = mapping.apply(mapping.parse_map(golden / "two.map", **kwargs), two)
        np.testing.assert_allclose(cg.xyz[0, 0], [expect] * 3, rtol=rtol)
    four = system.load(golden / "martini3" / "four.gro")
    for kwargs, expect, rtol in [({}, 2.5, 1e-6),
                                 ({"virtual_map_center": "mass", "itp_path": golden / "martini3" / "four.itp"}, 3.0, 1e-6),
                                 ({"virtual_map_center": "mass"}, 2.83337, 1e-2)]:
        cg = mapping.apply(mapping.parse_map(golden / "martini3" / "four.map", **kwargs), four)
        np.testing.assert_allclose(cg.xyz[0, 2], [expect] * 3, rtol=rtol)
    empty = mapping.apply(mapping.parse_map(golden / "water.map"), system.load(golden / "sugar.gro"))
    assert empty.xyz.shape == (1, 0, 3)


def test_circular_statistics_kats():
    # reference tests/test_util.py:176-195
    v = np.deg2rad(np.array([-175, 165], dtype=np.float32))
    assert stats.circular_mean(v) == pytest.approx(np.deg2rad(175.0), rel=1 / 500)
    w = np.deg2rad(np.array([165, 145], dtype=np.float32))
    assert stats.circular_mean(w) == pytest.approx(np.deg2rad(155.0), rel=1 / 500)
    assert stats.circular_variance(w) == pytest.approx(np.var(w), rel=1 / 500)
    assert stats.circular_variance(v) == pytest.approx(np.var(w), rel=1 / 500)


def test_graph_walk_kats():
    # reference tests/test_util.py:18-39 and SURVEY.md App. A4 (lipid bond graph)
    g = bonds.grow_chains
    pairs = [(0, 1), (1, 2), (2, 3)]
    assert sorted(g(pairs, pairs)) == [(0, 1, 2), (1, 2, 3)]
    ring = [(0, 1), (1, 2), (2, 3), (3, 0)]
    triplets = g(ring, ring)
    assert sorted(triplets) == [(0, 1, 2), (1, 0, 3), (1, 2, 3), (2, 3, 0)]
    assert sorted(g(triplets, ring)) == [(0, 1, 2, 3), (1, 0, 3, 2), (1, 2, 3, 0), (2, 1, 0, 3)]
    multi = [("a", "b"), ("b", "c"), ("c", "d"), ("d", "+a")]
    assert sorted(g(multi, multi)) == [("a", "b", "c"), ("b", "c", "d"), ("c", "d", "+a"), ("d", "+a", "+b")]


# ---------------------------------------------------------------------------------------------- triclinic wrap
# The reference holds no fixture in which a triclinic image shift happens: the oracle's restatement of
# mdtraj's general kernel is checked here against an exhaustive float64 image search written from the definition of
# the minimum image (tests/bruteforce.py), on boxes whose skew stays inside GROMACS' reduced-cell convention
# (|b_x| <= a_x / 2, |c_x| <= a_x / 2, |c_y| <= b_y / 2), where mdtraj's 27-image search is exact.
@pytest.mark.parametrize("skew", [(0.9, -1.1, 1.3), (1.45, 1.45, 1.45), (-1.4, 0.3, -1.2), (0.0, 1.2, 0.0)])
def test_oracle_triclinic_minimum_image_equals_exhaustive_search(skew):
    from oracle import geometry as ogeom
    from tests.bruteforce import min_image_distance, skewed_boxes
    rng = np.random.default_rng(11)
    n_frames, n_beads = 40, 60
    box = skewed_boxes(n_frames, (3.0, 3.1, 2.9), skew, seed=3)
    frac = rng.uniform(-0.2, 1.2, size=(n_frames, n_beads, 3))          # inside the cell and a little outside
    xyz = np.einsum("fbi,fij->fbj", frac, box.astype(np.float64)).astype(np.float32)
    pairs = np.array([(i, j) for i in range(0, n_beads, 3) for j in range(1, n_beads, 7) if i != j], dtype=np.int32)
    got = ogeom.distances(xyz, pairs, box).astype(np.float64)
    want = min_image_distance(xyz, pairs, box)
    assert not ogeom.is_orthorhombic(box)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    # the wrap really happens: many pairs are closer through an image than directly
    direct = np.linalg.norm(xyz[:, pairs[:, 1]].astype(np.float64) - xyz[:, pairs[:, 0]], axis=-1)
    assert (want < direct - 1e-3).mean() > 0.3


def test_oracle_triclinic_beyond_the_reduced_cell_is_documented():
    """Outside GROMACS' convention (skew of more than half a cell) mdtraj first reduces the box vectors; the
    restatement keeps that step, and the result still equals the exhaustive search."""
    from oracle import geometry as ogeom
    from tests.bruteforce import min_image_distance, skewed_boxes
    rng = np.random.default_rng(12)
    box = skewed_boxes(10, (3.0, 3.0, 3.0), (2.2, -1.9, 2.6), seed=4)
    frac = rng.uniform(0, 1, size=(10, 30, 3))
    xyz = np.einsum("fbi,fij->fbj", frac, box.astype(np.float64)).astype(np.float32)
    pairs = np.array([(i, (i * 7 + 3) % 30) for i in range(30) if i != (i * 7 + 3) % 30], dtype=np.int32)
    np.testing.assert_allclose(ogeom.distances(xyz, pairs, box).astype(np.float64),
                               min_image_distance(xyz, pairs, box, reach=4), rtol=2e-6, atol=2e-6)

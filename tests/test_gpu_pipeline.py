"""GPU parity of the fused mapping + measurement pass (``map_and_measure`` -> ``cgb_map_measure``).

The fused kernel must give what the two separate stages give -- the reference's call sequence
``Mapping.apply`` then ``BondSet.apply`` (pycgtool/__main__.py:69-112) -- and therefore what the oracle gives:
beads bit-exact, lengths bit-exact, angles / dihedrals identical to the separate kernels (same device code),
statistics and fitted parameters within 1e-4 of the oracle.
"""

import numpy as np
import pytest

from oracle import bonds as obonds, mapping as omap, system as osystem
from tests.helpers import oracle_system
from tests.test_gpu_bonds import _assert_params, _assert_values

pytestmark = pytest.mark.gpu


def _both_ways(sysm, n_frames, options, box=True, skew=None, histogram=True, store_values=True):
    """(fused BondSet, fused CG frame, separate BondSet, separate CG frame, oracle terms, oracle CG system)."""
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure, pipeline
    xyz = sysm.coordinates_host(n_frames)
    bx = sysm.box_vectors(n_frames) if box else None
    if skew is not None:
        bx[:, 1, 0], bx[:, 2, 0], bx[:, 2, 1] = skew
    map_path, bnd_path = sysm.write_files()
    mapper = Mapping(map_path, options)
    frame = Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=bx)
    sep = BondSet(bnd_path, options)
    cg_sep = mapper.apply(frame)
    sep.apply(cg_sep, histogram=histogram, store_values=store_values)
    sep.boltzmann_invert()
    fused = BondSet(bnd_path, options)
    cg_fused = map_and_measure(mapper, fused, frame, histogram=histogram, store_values=store_values, fused=True)
    assert pipeline._fused_plan(mapper, fused, mapper._plan_for(frame._topology)) is not None, "fused plan not used"
    fused.boltzmann_invert()
    ocg = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, bx))
    terms = obonds.parse_bnd(bnd_path, generate_angles=options.generate_angles,
                             generate_dihedrals=options.generate_dihedrals)
    obonds.measure(terms, ocg)
    obonds.invert(terms)
    return fused, cg_fused, sep, cg_sep, terms, ocg


def _check(fused, cg_fused, sep, cg_sep, terms, ocg, values=True):
    assert np.array_equal(cg_fused._trajectory.xyz, ocg.xyz), "fused beads differ from the oracle"
    assert np.array_equal(cg_fused._trajectory.xyz, cg_sep._trajectory.xyz)
    for mol in terms:
        assert len(fused[mol]) == len(terms[mol])
        for fb, sb, term in zip(fused[mol], sep[mol], terms[mol]):
            if values:
                assert np.array_equal(np.asarray(fb.values), np.asarray(sb.values), equal_nan=True)
                _assert_values(fb, term, ocg)
            _assert_params(fb, term)
            if sb.eqm is not None:
                assert fb.eqm == pytest.approx(sb.eqm, rel=1e-6, abs=1e-7)
                assert fb.fconst == pytest.approx(sb.fconst, rel=1e-5)
            cf, ef = fb.histogram
            cs, es = sb.histogram
            assert np.array_equal(cf, cs) and np.array_equal(ef, es)


@pytest.mark.parametrize("n_frames", [1, 3, 64, 131])
def test_membrane_with_solvent_fused_equals_separate_and_oracle(options, n_frames):
    from workloads import synth
    sysm = synth.membrane(40, 60, seed=17)
    sysm.box = np.array([4.0, 4.2, 3.9])
    sysm.sigma = 0.12
    _check(*_both_ways(sysm, n_frames, options))


def test_membrane_with_dihedrals_fused(options):
    from workloads import synth
    sysm = synth.membrane(25, 10, seed=5)
    sysm.box = np.array([3.5, 3.6, 3.7])
    sysm.sigma = 0.12
    options.generate_dihedrals = True
    _check(*_both_ways(sysm, 77, options))


def test_small_molecule_fused(options):
    from workloads import synth
    _check(*_both_ways(synth.small_molecule(), 2050, options))


def test_fused_without_box_and_triclinic(options):
    from workloads import synth
    sysm = synth.membrane(12, 7, seed=4)
    sysm.box = np.array([3.0, 3.0, 3.0])
    sysm.sigma = 0.1
    _check(*_both_ways(sysm, 21, options, box=False))
    _check(*_both_ways(sysm, 21, options, skew=(0.9, -1.1, 1.3)))


def test_fused_stats_only_matches_stored(options):
    """stats-only (no values written): same statistics and histograms; ``bond.values`` re-measured on demand."""
    from workloads import synth
    sysm = synth.membrane(30, 0, seed=9)
    sysm.box = np.array([3.8, 3.8, 3.8])
    sysm.sigma = 0.12
    fused, cg_fused, sep, cg_sep, terms, ocg = _both_ways(sysm, 50, options, store_values=False)
    assert fused._state.values is None
    _check(fused, cg_fused, sep, cg_sep, terms, ocg, values=False)
    for fb, term in zip(fused["DOPC"], terms["DOPC"]):
        _assert_values(fb, term, ocg)           # lazily re-measured from the CG frame


def test_sugar_fused(golden, options):
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure
    from oracle import fileio
    import json
    frame = Frame(golden / "sugar.gro", golden / "sugar.xtc")
    bonds = BondSet(golden / "sugar.bnd", options)
    cg = map_and_measure(Mapping(golden / "sugar.map", options), bonds, frame, fused=True)
    bonds.boltzmann_invert()
    gold, _, _ = fileio.read_xtc(golden / "sugar_out.xtc")
    assert np.array_equal(cg._trajectory.xyz, gold)
    stored = json.loads((golden / "sugar_params.json").read_text())
    for bond, row in zip(bonds["ALLA"], stored):
        assert bond.eqm == pytest.approx(row["eqm"], rel=1e-4, abs=2e-6)
        assert bond.fconst == pytest.approx(row["fconst"], rel=1e-4)


def test_virtual_beads_through_map_and_measure(golden, options, tmp_path):
    """martini3 naphthalene (a virtual site built from the real beads): the .itp still matches the reference's file
    (reference tests/test_bondset.py:210-235) whichever way ``map_and_measure`` takes."""
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure
    from pycgtool_b200.util import cmp_file_whitespace_float
    options.generate_angles = False
    measure = BondSet(golden / "martini3" / "naphthalene.bnd", options)
    mapping = Mapping(golden / "martini3" / "naphthalene.map", options)
    frame = Frame(golden / "martini3" / "naphthalene.gro")
    cg = map_and_measure(mapping, measure, frame, fused=True)
    assert np.array_equal(cg._trajectory.xyz, mapping.apply(frame)._trajectory.xyz)
    measure.boltzmann_invert()
    measure.write_itp(tmp_path / "naphthalene_out.itp", mapping)
    assert cmp_file_whitespace_float(tmp_path / "naphthalene_out.itp", golden / "martini3" / "naphthalene_out.itp",
                                     rtol=0.005, verbose=True)


def test_unsupported_systems_fall_back_to_two_stages(options):
    """A molecule of 60 beads does not fit a fused tile (48 beads): ``map_and_measure`` runs K1 + K2, same results."""
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure, pipeline
    from workloads import synth
    sysm = synth.small_molecule(n_beads=60)
    xyz, box = sysm.coordinates_host(40), sysm.box_vectors(40)
    map_path, bnd_path = sysm.write_files()
    mapper = Mapping(map_path, options)
    frame = Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box)
    a, b = BondSet(bnd_path, options), BondSet(bnd_path, options)
    cg_a = map_and_measure(mapper, a, frame, fused=True)
    assert pipeline._fused_plan(mapper, a, mapper._plan_for(frame._topology)) is None
    cg_b = mapper.apply(frame)
    b.apply(cg_b)
    assert np.array_equal(cg_a._trajectory.xyz, cg_b._trajectory.xyz)
    a.boltzmann_invert()
    b.boltzmann_invert()
    for x, y in zip(a["MOL"], b["MOL"]):
        assert np.array_equal(np.asarray(x.values), np.asarray(y.values))
        # float64 sums of different CTAs meet in global atomics: the last bits depend on their order
        assert x.eqm == pytest.approx(y.eqm, rel=1e-9) and x.fconst == pytest.approx(y.fconst, rel=1e-9)


def test_preallocated_output_buffer_is_used(options):
    """``out=``: a pipeline that maps block after block writes the beads into one buffer (both paths)."""
    import torch
    from pycgtool_b200 import BondSet, Frame, Mapping, map_and_measure
    from workloads import synth
    sysm = synth.membrane(20, 10, seed=8)
    sysm.box = np.array([3.5, 3.5, 3.5])
    sysm.sigma = 0.12
    xyz, box = sysm.coordinates_host(18), sysm.box_vectors(18)
    map_path, bnd_path = sysm.write_files()
    mapper = Mapping(map_path, options)
    frame = Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box)
    want = mapper.apply(frame)._trajectory.xyz
    n_beads = want.shape[1]
    for fused in (False, True):
        buf = torch.full((18, n_beads, 3), float("nan"), dtype=torch.float32, device="cuda")
        bonds = BondSet(bnd_path, options)
        cg = map_and_measure(mapper, bonds, frame, fused=fused, out=buf)
        assert cg._trajectory.device_xyz().data_ptr() == buf.data_ptr()
        assert np.array_equal(buf.cpu().numpy(), want)
        bonds.boltzmann_invert()
        assert all(b.eqm is not None for b in bonds["DOPC"])
    with pytest.raises(ValueError):
        mapper.apply(frame, out=torch.empty((3, n_beads, 3), dtype=torch.float32, device="cuda"))

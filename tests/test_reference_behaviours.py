"""Host-side behaviours of the hot path's callers, mirroring the reference's own unit tests
(tests/test_util.py, tests/test_functionalforms.py, tests/test_frame.py): same inputs, same expected
results, through this package's modules.  No GPU involved."""

import filecmp
import os
import pathlib

import numpy as np
import pytest

from pycgtool_b200 import Frame, util
from pycgtool_b200.frame import NonMatchingSystemError, UnsupportedFormatException
from pycgtool_b200.functionalforms import FunctionalForm, get_functional_forms


# ------------------------------------------------------------------ util (reference tests/test_util.py)
def test_extend_graph_chain_triplets_and_quadruplets():
    pairs = [(0, 1), (1, 2), (2, 3)]
    assert sorted(util.extend_graph_chain(pairs, pairs)) == [(0, 1, 2), (1, 2, 3)]
    ring = [(0, 1), (1, 2), (2, 3), (3, 0)]
    assert sorted(util.extend_graph_chain(ring, ring)) == [(0, 1, 2), (1, 0, 3), (1, 2, 3), (2, 3, 0)]
    multires = [("a", "b"), ("b", "c"), ("c", "d"), ("d", "+a")]
    assert sorted(util.extend_graph_chain(multires, multires)) == [
        ("a", "b", "c"), ("b", "c", "d"), ("c", "d", "+a"), ("d", "+a", "+b")]
    triplets = util.extend_graph_chain(pairs, pairs)
    assert sorted(util.extend_graph_chain(triplets, pairs)) == [(0, 1, 2, 3)]
    triplets = util.extend_graph_chain(ring, ring)
    assert sorted(util.extend_graph_chain(triplets, ring)) == [(0, 1, 2, 3), (1, 0, 3, 2), (1, 2, 3, 0), (2, 1, 0, 3)]


def test_dir_up():
    path = os.path.realpath(__file__)
    assert util.dir_up(path, 0) == path
    assert util.dir_up(path) == os.path.dirname(path)
    assert util.dir_up(path, 2) == os.path.dirname(os.path.dirname(path))


@pytest.mark.parametrize("make", [pathlib.Path.touch, pathlib.Path.mkdir])
def test_backup_file_and_directory(tmp_path, make):
    path = tmp_path / "testfile"
    make(path)
    assert util.backup_file(path) == tmp_path / "#testfile.1#"
    assert not path.exists() and (tmp_path / "#testfile.1#").exists()
    make(path)
    assert util.backup_file(path) == tmp_path / "#testfile.2#"
    assert (tmp_path / "#testfile.1#").exists() and (tmp_path / "#testfile.2#").exists()


def test_sliding_window():
    assert list(util.sliding([0, 1, 2, 3, 4])) == [(None, 0, 1), (0, 1, 2), (1, 2, 3), (2, 3, 4), (3, 4, None)]
    with pytest.raises(ValueError):
        list(util.sliding([]))


def test_transpose_and_sample():
    rows = [(1, 2), (3, 4), (5, 6)]
    assert util.transpose_and_sample(rows, None) == [(1, 3, 5), (2, 4, 6)]
    sampled = util.transpose_and_sample(rows, n=1)
    assert len(sampled) == 1 and sampled[0] in [(1, 3, 5), (2, 4, 6)]


def test_cmp_whitespace_float():
    assert util.cmp_whitespace_float(["Hello World"], ["Hello World"])
    assert not util.cmp_whitespace_float(["Hello World"], ["Hello PyCGTOOL"])
    ref = ["8.3 1.00 -3 0"]
    assert util.cmp_whitespace_float(ref, ["8.3 1.00 -3 0.0"])
    assert not util.cmp_whitespace_float(ref, ["8.3 1.01 -3 0"])
    assert util.cmp_whitespace_float(ref, ["8.3 1.01 -3 0"], rtol=0.1)
    assert util.cmp_whitespace_float(ref, ["8.3 1.00 -2.999 0"])
    assert util.cmp_whitespace_float(ref, ["8.3 1.00 -3.001 0"])
    assert not util.cmp_whitespace_float(["1"], ["1", "2"])


def test_circular_statistics():
    values = np.deg2rad(np.array([-175, 165], dtype=np.float32))
    assert util.circular_mean(values) == pytest.approx(np.deg2rad(175.0), rel=1 / 500)
    plain = np.deg2rad(np.array([165, 145], dtype=np.float32))
    assert util.circular_mean(plain) == pytest.approx(np.deg2rad(155.0), rel=1 / 500)
    var = np.var(plain)
    assert util.circular_variance(plain) == pytest.approx(var, rel=1 / 500)
    assert util.circular_variance(values) == pytest.approx(var, rel=1 / 500)


def test_file_write_lines(tmp_path):
    empty = tmp_path / "empty"
    util.file_write_lines(empty)
    assert empty.read_text() == ""
    lines = ["a", "b c", "  ", "123 d", "", " %-+# "]
    util.file_write_lines(tmp_path / "list", lines)
    assert (tmp_path / "list").read_text().splitlines() == lines
    util.file_write_lines(tmp_path / "append", lines[:2])
    util.file_write_lines(tmp_path / "append", lines[2:], append=True)
    assert (tmp_path / "append").read_text().splitlines() == lines


def test_compare_trajectories(golden):
    assert util.compare_trajectories(golden / "sugar.gro", golden / "sugar.gro")
    with pytest.raises(ValueError):
        util.compare_trajectories(golden / "sugar.gro", golden / "water.gro")
    assert util.compare_trajectories(golden / "sugar.xtc", golden / "sugar.xtc", topology_file=golden / "sugar.gro")
    with pytest.raises(ValueError):
        util.compare_trajectories(golden / "sugar.xtc", golden / "water.xtc", topology_file=golden / "sugar.gro")


# ------------------------------------------------------------------ functional forms (tests/test_functionalforms.py)
def test_functional_forms_are_discovered_including_user_subclasses():
    funcs = get_functional_forms()
    assert funcs.Harmonic is not None and funcs.CosHarmonic is not None

    class TestFunc(FunctionalForm):
        gromacs_type_ids = (None, None, None)

        @staticmethod
        def eqm(values, temp):
            return "TestResultEqm"

        @staticmethod
        def fconst(values, temp):
            return "TestResultFconst"

    test_func = get_functional_forms()["TestFunc"].value
    assert test_func.eqm(None, None) == "TestResultEqm" and test_func.fconst(None, None) == "TestResultFconst"


def test_injected_mean_and_variance_functions():
    funcs = get_functional_forms()
    vals = [0.25 * np.pi, 1.75 * np.pi]
    assert funcs["Harmonic"].value().eqm(vals, None) == pytest.approx(np.pi)
    assert funcs["Harmonic"].value(mean_func=util.circular_mean).eqm(vals, None) == pytest.approx(0, abs=1e-6)
    func = funcs["Harmonic"].value(variance_func=util.circular_variance)
    assert func.fconst(vals, 300) == pytest.approx(func.fconst([0, 0.5 * np.pi], 300), rel=1e-12)


# ------------------------------------------------------------------ Frame (tests/test_frame.py)
WATER_ATOM0 = np.array([
    [1.176, 1.1520001, 1.5860001], [1.1220001, 1.13, 1.534], [1.0580001, 1.1620001, 1.462],
    [0.91600007, 1.276, 1.5580001], [0.73200005, 1.1240001, 1.286], [0.64000005, 1.2160001, 1.258],
    [0.632, 1.312, 1.2520001], [0.606, 1.284, 1.246], [0.582, 1.312, 1.1600001], [0.68200004, 1.22, 1.25],
    [0.69600004, 1.33, 1.21]], dtype=np.float32)
WATER_BOX = np.array([1.9052, 1.9032527, 1.9040661, 1.896811, 1.8985983, 1.9033976, 1.8904614, 1.9013108,
                      1.8946321, 1.898091, 1.898684], dtype=np.float32)


def _check_water_topology(frame):
    assert frame.natoms == 663 and len(list(frame.residues)) == 221
    assert frame.residue(0).n_atoms == 3


def _check_water_frame(frame, check_box=True):
    _check_water_topology(frame)
    np.testing.assert_allclose(frame.atom(0).coords, [[0.696, 1.330, 1.211]], rtol=1e-6)
    if check_box:
        np.testing.assert_allclose(frame.unitcell_lengths, [[1.89868, 1.89868, 1.89868]], rtol=1e-4)


def test_frame_create_and_add_residue():
    frame = Frame()
    residue = frame.add_residue("TEST_RESIDUE")
    assert residue == frame.residue(0) and residue.name == "TEST_RESIDUE"


def test_frame_reads_gro_and_pdb(golden):
    _check_water_frame(Frame(golden / "water.gro"))
    _check_water_frame(Frame(golden / "water.pdb"))
    _check_water_frame(Frame(golden / "water-nobox.pdb"), check_box=False)
    assert Frame(golden / "water-nobox.pdb").unitcell_lengths is None
    assert Frame(golden / "polyethene.gro").unitcell_lengths is None      # zero box
    with pytest.raises(UnsupportedFormatException):
        Frame(golden / "dppc.map")


def test_frame_reads_and_writes_trajectories(golden, tmp_path):
    frame = Frame(golden / "water.gro", golden / "water.xtc")     # pathlib.Path arguments are accepted
    assert frame.n_frames == 11
    _check_water_topology(frame)
    np.testing.assert_allclose(frame.atom(0).coords, WATER_ATOM0, rtol=1e-6)
    np.testing.assert_allclose(frame.unitcell_lengths, np.repeat(WATER_BOX[:, None], 3, axis=1), rtol=1e-4)
    frame.save(tmp_path / "water.xtc")
    assert util.compare_trajectories(golden / "water.xtc", tmp_path / "water.xtc", topology_file=golden / "water.gro")
    single = Frame(golden / "water.gro")
    single.save(tmp_path / "water-out.gro")
    assert filecmp.cmp(golden / "water.gro", tmp_path / "water-out.gro", shallow=False)
    with pytest.raises(NonMatchingSystemError):
        Frame(golden / "water.gro", golden / "sugar.xtc")

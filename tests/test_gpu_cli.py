"""End-to-end workflows through the command line front end, mirroring the reference's
tests/test_pycgtool.py (map only, full run, no trajectory, measure only, force-field directory)."""

import pathlib

import numpy as np
import pytest

from oracle import fileio

pytestmark = pytest.mark.gpu


def _run(golden, tmp_path, *extra, traj=True, mapping=True, bnd=True):
    from pycgtool_b200.__main__ import main
    argv = [str(golden / "sugar.gro")]
    if traj:
        argv.append(str(golden / "sugar.xtc"))
    if mapping:
        argv += ["-m", str(golden / "sugar.map")]
    if bnd:
        argv += ["-b", str(golden / "sugar.bnd")]
    argv += ["--out-dir", str(tmp_path), "--log-level", "ERROR", *extra]
    return main(argv)


def test_map_only_writes_the_reference_trajectory(golden, tmp_path):
    # reference tests/test_pycgtool.py:70-92 (compare_trajectories, rtol 1e-3); here: identical floats
    _run(golden, tmp_path, "--output-xtc", bnd=False)
    mine, _, box = fileio.read_xtc(tmp_path / "out.xtc")
    gold, _, gbox = fileio.read_xtc(golden / "sugar_out.xtc")
    assert np.array_equal(mine, gold)
    np.testing.assert_allclose(box, gbox, rtol=1e-6)


def test_full_run_itp_and_gro(golden, tmp_path):
    # reference tests/test_pycgtool.py:94-118
    from pycgtool_b200 import formats
    from pycgtool_b200.util import cmp_file_whitespace_float
    _run(golden, tmp_path)
    assert cmp_file_whitespace_float(tmp_path / "out.itp", golden / "sugar_out.itp", rtol=0.001, verbose=True)
    mine, ref = formats.read_gro(tmp_path / "out.gro"), formats.read_gro(golden / "sugar_out.gro")
    assert mine.topology == ref.topology
    np.testing.assert_allclose(mine.xyz, ref.xyz, rtol=1e-3)


def test_no_trajectory_maps_the_topology_frame(golden, tmp_path):
    # reference tests/test_pycgtool.py:120-143: single frame -> no inversion, no itp
    _run(golden, tmp_path, traj=False)
    assert (tmp_path / "out.gro").exists() and not (tmp_path / "out.itp").exists()


def test_measure_only_dumps_samples(golden, tmp_path):
    # reference tests/test_pycgtool.py:145-173: bondset without mapping measures the input frame itself
    from pycgtool_b200.__main__ import main
    from pycgtool_b200.util import cmp_file_whitespace_float
    # the .bnd names (C1, C2, ...) are resolved against the atomistic topology, as in the reference test
    main([str(golden / "sugar.gro"), "-b", str(golden / "sugar.bnd"), "--out-dir", str(tmp_path),
          "--log-level", "ERROR"])
    assert not (tmp_path / "out.itp").exists()
    for kind in ("length", "angle", "dihedral"):
        assert cmp_file_whitespace_float(golden / f"ALLA_{kind}_one.dat", tmp_path / f"ALLA_{kind}.dat",
                                         rtol=1e-4, verbose=True)


def test_forcefield_rtp_matches_reference_output(golden, tmp_path):
    """``out.rtp`` of the reference is produced by its current (circular-statistics) algorithm, so every
    printed number must agree within the north star's 1e-4 (the reference's own test compares the file
    with itself, tests/test_pycgtool.py:198)."""
    from pycgtool_b200.util import cmp_file_whitespace_float
    _run(golden, tmp_path, "--output-forcefield")
    ff = tmp_path / "ffout.ff"
    assert cmp_file_whitespace_float(golden / "ffout.ff" / "out.rtp", ff / "out.rtp", rtol=1e-4, verbose=True)
    assert (ff / "out.r2b").read_text() == (golden / "ffout.ff" / "out.r2b").read_text()
    assert (ff / "forcefield.doc").read_text().strip() == "PyCGTOOL produced MARTINI force field - out"

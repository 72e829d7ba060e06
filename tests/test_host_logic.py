"""CPU tests of the host side: C-ABI surface, plan compilers, parsers, graph walk.

No compute kernels are launched here.  "Assignment" (every integer table the
kernels consume) must be bit-exact with the reference's name lookups, which the
oracle restates with the reference's own loops.
"""

import ctypes
import pathlib
import re

import numpy as np
import pytest

from oracle import bonds as obonds, mapping as omap, system as osystem
from tests.helpers import oracle_system

ROOT = pathlib.Path(__file__).resolve().parent.parent


# ----------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from pycgtool_b200 import _lib
    header = (ROOT / "include" / "pycgtool_b200.h").read_text()
    declared = set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", header))
    assert {"cgb_map_tiles", "cgb_measure", "cgb_measure_fused", "cgb_measure_plan_build", "cgb_boltzmann",
            "cgb_map_plan_build", "cgb_allreduce_partials"} <= declared
    lib = ctypes.CDLL(str(_lib.LIBRARY_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == declared
    loaded = _lib.load()
    assert loaded.cgb_abi_version() == _lib.ABI_VERSION == 3
    assert b"invalid argument" in loaded.cgb_strerror(-1)


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable without a device."""
    from pycgtool_b200 import _lib
    lib = _lib.load()
    assert lib.cgb_map_tiles(None, None, 4, 10, 2, None, 1, None, None, 8, 8, None, None) == -1
    assert lib.cgb_map_tiles(None, None, 0, 10, 2, None, 1, None, None, 8, 8, None, None) == 0   # no frames: no-op
    assert lib.cgb_measure(None, None, 1, -1, 2, None, None, None, 0, 0, 0, 0, None, None, 0, None, None, None, None,
                           0, None) == -1
    assert lib.cgb_measure_fused(None, None, 1, 5, 2, None, 1, None, None, None, None, 0, 8, 8, 1, 1, 1, None, None, 0,
                                 None, None, None, None, 0, None, None) == -1
    # frame folding: the folded frame must be a whole number of copies
    assert lib.cgb_measure_fused(None, None, 1, 5, 7, None, 1, None, None, None, None, 0, 8, 8, 1, 1, 2, None, None, 0,
                                 None, None, None, None, 0, None, None) == -1
    assert lib.cgb_boltzmann(None, None, None, None, 310.0, 3, None, None, None, 0, 0, None, None) == -1
    assert lib.cgb_allreduce_partials(None, None, 4, None) == -1
    assert lib.cgb_allreduce_partials(None, None, 0, None) == 0       # nothing to reduce: no-op
    bad = _lib.MapTuning(stage_bytes=-5)
    # tuning is a per-call argument (no process-wide state); a bad one is rejected before any launch
    assert lib.cgb_map_tiles_ex(None, None, 0, 10, 0, 2, None, 1, None, None, 8, 8, None, _lib.tuning_ptr(bad), None) == -1
    assert lib.cgb_map_tiles_ex(None, None, 0, 10, 0, 2, None, 1, None, None, 8, 8, None, None, None) == 0
    with pytest.raises(RuntimeError):
        _lib.check(-3, "demo")


def _decode_tiles(tiles, entries, n_beads):
    """Rebuild the CSR assignment from a tile plan (inverse of cgb_map_plan_build)."""
    tiles = tiles.reshape(-1, 8)
    ent = entries.reshape(-1, 2)
    rows = {}
    for t in tiles:
        atom0, span, bead0, nbeads = (int(x) for x in t[:4])
        ent0 = int(np.array(t[4:6], dtype=np.int32).view(np.int64)[0])
        nslots = int(t[6])
        for j in range(nbeads):
            cnt = int(ent[ent0 + j, 1])
            if cnt == 0:
                continue
            atoms, weights = [], []
            for s in range(cnt):
                off3, bits = ent[ent0 + s * nbeads + j]
                assert off3 % 3 == 0 and 0 <= off3 // 3 < span
                atoms.append(atom0 + off3 // 3)
                weights.append(np.array([bits], dtype=np.int32).view(np.float32)[0])
            for s in range(cnt, nslots):          # padding repeats the reference atom with weight 0
                off3, bits = ent[ent0 + s * nbeads + j]
                assert off3 == ent[ent0 + j, 0] and bits == 0
            assert bead0 + j not in rows
            rows[bead0 + j] = (atoms, weights)
    return rows


@pytest.mark.parametrize("max_span,max_beads", [(512, 512), (140, 7), (20000, 3)])
def test_tile_plan_encodes_the_same_assignment(options, max_span, max_beads):
    from pycgtool_b200 import Mapping, _lib
    from workloads import synth
    lib = _lib.load()
    sysm = synth.membrane(5, 33, seed=2)
    map_path, _ = sysm.write_files()
    plan = Mapping(map_path, options).compile_plan(sysm.topology)
    n_tiles, n_ent = ctypes.c_int64(), ctypes.c_int64()
    ms, me = ctypes.c_int32(), ctypes.c_int32()
    args = (_lib.ptr(plan.bead_ptr), _lib.ptr(plan.index), None, plan.n_beads, plan.n_atoms, max_span, max_beads)
    _lib.check(lib.cgb_map_plan_size(*args, ctypes.byref(n_tiles), ctypes.byref(n_ent), ctypes.byref(ms),
                                     ctypes.byref(me)))
    tiles = np.zeros(n_tiles.value * 8, dtype=np.int32)
    entries = np.zeros(n_ent.value * 2, dtype=np.int32)
    _lib.check(lib.cgb_map_plan_build(_lib.ptr(plan.bead_ptr), _lib.ptr(plan.index), _lib.ptr(plan.weights), 0,
                                      None, plan.n_beads, plan.n_atoms, max_span, max_beads,
                                      _lib.ptr(tiles), _lib.ptr(entries), None))
    t = tiles.reshape(-1, 8)
    assert (t[:, 1] <= max_span).all() and (t[:, 3] <= max_beads).all() and ms.value == t[:, 1].max()
    rows = _decode_tiles(tiles, entries, plan.n_beads)
    assert sorted(rows) == list(range(plan.n_beads))
    for b in range(plan.n_beads):
        lo, hi = plan.bead_ptr[b], plan.bead_ptr[b + 1]
        assert rows[b][0] == list(plan.index[lo:hi])
        assert rows[b][1][1:] == list(plan.weights[lo + 1:hi])


def test_plan_compiler_rejects_bad_input():
    from pycgtool_b200 import _lib
    lib = _lib.load()
    ptr = np.array([0, 2], dtype=np.int32)
    out = [ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()]
    refs = [ctypes.byref(o) for o in out]
    idx = np.array([0, 9], dtype=np.int32)
    assert lib.cgb_map_plan_size(_lib.ptr(ptr), _lib.ptr(idx), None, 1, 5, 512, 512, *refs) == -2   # index >= n_atoms
    idx = np.array([0, 4], dtype=np.int32)
    assert lib.cgb_map_plan_size(_lib.ptr(ptr), _lib.ptr(idx), None, 1, 5, 3, 512, *refs) == -3     # wider than a tile
    assert lib.cgb_map_plan_size(_lib.ptr(ptr), _lib.ptr(idx), None, 1, 5, 512, 512, *refs) == 0
    assert out[0].value == 1 and out[1].value == 2 and out[2].value == 5


# ----------------------------------------------------------------------------- assignment parity
@pytest.mark.parametrize("gro,mp,kwargs", [
    ("sugar.gro", "sugar.map", {}), ("water.gro", "water.map", {}), ("water.gro", "water_double_weight.map", {}),
    ("two.gro", "two.map", {"map_center": "mass", "itp": "two.itp"}),
    ("martini3/four.gro", "martini3/four.map", {"virtual_map_center": "mass", "itp": "martini3/four.itp"}),
    ("martini3/naphthalene.gro", "martini3/naphthalene.map", {}), ("sugar.gro", "water.map", {})])
def test_mapping_assignment_bit_exact(golden, options, gro, mp, kwargs):
    from pycgtool_b200 import Frame, Mapping
    itp = kwargs.pop("itp", None)
    for key, val in kwargs.items():
        setattr(options, key, val)
    frame = Frame(golden / gro)
    mapper = Mapping(golden / mp, options, itp_filename=golden / itp if itp else None)
    plan = mapper.compile_plan(frame._topology)
    omols = omap.parse_map(golden / mp, options.map_center, options.virtual_map_center, golden / itp if itp else None)
    ptr, idx, virt = omap.assignment(omols, osystem.load(golden / gro))
    assert np.array_equal(plan.bead_ptr, ptr) and np.array_equal(plan.index, idx)
    assert np.array_equal(plan.is_virtual.astype(bool), virt)
    # weights: same values and the same dtype (float64 only for ITP masses, mapping.py:224-227)
    sysm = osystem.load(golden / gro)
    expect_w = [w for res in omap.mapped_residues(omols, sysm) for bead in omols[res.name] for w in bead.weights]
    combined = np.where(np.repeat(plan.is_virtual.astype(bool), np.diff(plan.bead_ptr)), plan.vweights, plan.weights)
    assert np.array_equal(combined.astype(np.float64), np.array(expect_w, dtype=np.float64))
    # the CG topology the reference builds in _cg_frame_setup (mapping.py:359-389)
    cg = plan.cg_topology
    expect = [(res.name, [b.name for b in omols[res.name]]) for res in omap.mapped_residues(omols, sysm)]
    got = [(cg.residue(r).name, [a.name for a in cg.residue(r).atoms]) for r in range(cg.n_residues)]
    assert got == expect


def test_mapping_assignment_synthetic_with_unmapped_residues(options):
    from pycgtool_b200 import Mapping
    from workloads import synth
    sysm = synth.s5(scale=0.002)
    map_path, _ = sysm.write_files()
    text = map_path.read_text().replace("[ CHOL ]", "[ NOTCHOL ]")     # CHOL residues become unmapped
    map_path.write_text(text)
    plan = Mapping(map_path, options).compile_plan(sysm.topology)
    osys = oracle_system(sysm.topology, np.zeros((1, sysm.n_atoms, 3), dtype=np.float32))
    ptr, idx, virt = omap.assignment(omap.parse_map(map_path), osys)
    assert np.array_equal(plan.bead_ptr, ptr) and np.array_equal(plan.index, idx) and not virt.any()


@pytest.mark.parametrize("gro,bnd", [("sugar-cg.gro", "sugar.bnd"), ("polyethene.gro", "polyethene.bnd"),
                                     ("pbcpolyethene.gro", "polyethene.bnd")])
def test_bond_assignment_bit_exact(golden, options, gro, bnd):
    from pycgtool_b200 import BondSet, Frame
    frame = Frame(golden / gro)
    bondset = BondSet(golden / bnd, options)
    plan = bondset.compile_plan(frame._topology)
    terms = obonds.parse_bnd(golden / bnd)
    sysm = osystem.load(golden / gro)
    bond_id = 0
    for mol, tlist in terms.items():
        assert [t.atoms for t in tlist] == [tuple(b.atoms) for b in bondset[mol]]
        for term in tlist:
            idx = obonds.resolve_indices(term, mol, sysm)
            assert np.array_equal(plan.bond_indices[bond_id], idx)
            cols = plan.bond_columns[bond_id]
            k = len(term.atoms)
            assert np.array_equal(plan.col_beads[cols][:, :k], idx)
            assert (plan.col_bond[cols] == bond_id).all()
            bond_id += 1
    assert plan.n_bonds == bond_id and sum(plan.counts) == len(plan.col_bond)
    # kinds are laid out lengths | angles | dihedrals
    n2, n3, n4 = plan.counts
    kinds = plan.natoms[plan.col_bond]
    assert (kinds[:n2] == 2).all() and (kinds[n2:n2 + n3] == 3).all() and (kinds[n2 + n3:] == 4).all()
    # the fused kernel's plan measures every column once, from its own beads, at the position bond_cols names
    from tests.plan_decode import check_measure_plan
    decoded = check_measure_plan(plan)
    for b, cols in enumerate(plan.bond_cols):
        for inst, pos in enumerate(cols):
            kind, beads, bond = decoded[int(pos)]
            assert bond == b and (beads == tuple(plan.bond_indices[b][inst]) or kind == 2)


@pytest.mark.parametrize("system,dihedrals", [("small", False), ("membrane", False), ("membrane", True),
                                              ("s5", False), ("chain", False)])
def test_measure_plan_covers_every_column(options, system, dihedrals, tmp_path):
    """``cgb_measure_plan_build`` (host code in the library): groups of <= 64 shared bead pairs, tiles of <= 7 groups
    inside one bead range, wide columns left to the gather kernel -- decoded back and compared column by column."""
    from pycgtool_b200 import BondSet, Mapping, _lib
    from pycgtool_b200.topology import Topology
    from workloads import synth
    from tests.plan_decode import check_measure_plan
    options.generate_dihedrals = dihedrals
    if system == "chain":
        names = [f"B{i}" for i in range(4000)]
        cg_top = Topology.from_residue_lists([("MOL", names)])
        lines = ["[ MOL ]"] + [f"B{i} B{i + 1}" for i in range(0, 3999, 7)] + ["B0 B3999", "B1 B3998"]
        lines += ["B0 B2000 B3999", "B5 B6 B7"] + ["B0 B1 B3998 B3999", "B10 B11 B12 B13"]
        bnd_path = tmp_path / "long.bnd"
        bnd_path.write_text("\n".join(lines) + "\n")
    else:
        sysm = {"small": synth.small_molecule, "membrane": lambda: synth.membrane(70, 40, seed=23),
                "s5": lambda: synth.s5(scale=0.02)}[system]()
        map_path, bnd_path = sysm.write_files()
        cg_top = Mapping(map_path, options).compile_plan(sysm.topology).cg_topology
    bondset = BondSet(bnd_path, options)
    plan = bondset.compile_plan(cg_top)
    check_measure_plan(plan)
    assert (plan.tiles["ngroups"] <= _lib.MEAS_WARPS).all() and (plan.tiles["span"] <= 544).all()
    assert plan.shape == (_lib.SHAPE_EAD if plan.counts[2] else _lib.SHAPE_EEAA)
    per_molecule = max(len(bonds) for bonds in bondset._molecules.values())
    assert plan.max_slots <= max(per_molecule, 8)
    # what the plan asks of a CTA's shared memory fits: tables + histograms + scratch + records + two stages of a quad
    need = (128 + plan.max_slots * (22 + 4 * 128) + 2 * 897 + 896 * 24 + 7 * 4096 + 2 * (4 * (12 * plan.max_span + 76) + 320))
    assert need <= 112 * 1024
    if system == "chain":
        assert sum(plan.wide_counts) == 4 and plan.wide_counts == (2, 1, 1)     # the terms that join the two ends
    else:
        assert sum(plan.wide_counts) == 0
    # forced gather path: every column is "wide", in its original order
    bondset.staged = False
    gather = bondset.compile_plan(cg_top)
    assert len(gather.tiles) == 0 and gather.wide_counts == gather.counts
    assert np.array_equal(gather.col_out, np.arange(gather.n_cols))


def test_generated_angles_and_s5_term_count(options):
    from pycgtool_b200 import BondSet
    from workloads import synth
    sysm = synth.s5(scale=0.001)
    _, bnd_path = sysm.write_files()
    bondset = BondSet(bnd_path, options)
    # 3 twelve-bead lipids x (11 bonds + 11 angles) + one eight-bead lipid x (7 + 6) = 79 Bond objects
    assert sum(len(bondset[m]) for m in bondset) == 79
    angles = [b.atoms for b in bondset["DOPC"] if len(b.atoms) == 3]
    assert angles == [("NC3", "PO4", "GL1"), ("PO4", "GL1", "GL2"), ("PO4", "GL1", "C1A"), ("GL1", "GL2", "C1B"),
                      ("GL2", "GL1", "C1A"), ("GL1", "C1A", "C2A"), ("C1A", "C2A", "C3A"), ("C2A", "C3A", "C4A"),
                      ("GL2", "C1B", "C2B"), ("C1B", "C2B", "C3B"), ("C2B", "C3B", "C4B")]
    terms = obonds.parse_bnd(bnd_path)
    for mol in terms:
        assert [t.atoms for t in terms[mol]] == [tuple(b.atoms) for b in bondset[mol]]


def test_graph_walk_matches_oracle_on_random_graphs():
    import random
    from pycgtool_b200.util import extend_graph_chain
    random.seed(7)
    for _ in range(300):
        names = [f"B{i}" for i in range(random.randint(2, 9))]
        edges = set()
        for _ in range(random.randint(1, 14)):
            a, b = random.sample(names, 2)
            edges.add((a, "+" + b if random.random() < 0.2 else b))
        edges = list(edges)
        triples = extend_graph_chain(edges, edges)
        assert triples == obonds.grow_chains(edges, edges)
        assert extend_graph_chain(triples, edges) == obonds.grow_chains(triples, edges)


# ----------------------------------------------------------------------------- parsers / containers
def test_bondset_parsing(golden, options):
    # reference tests/test_bondset.py:58-62,87-92,237-239
    from pycgtool_b200 import BondSet
    measure = BondSet(golden / "sugar.bnd", options)
    assert len(measure) == 1 and "ALLA" in measure and len(measure["ALLA"]) == 18
    tri = BondSet(golden / "triangle.bnd", options)
    assert len(tri.get_bond_angles("TRI", exclude_triangle=False)) == 3
    assert len(tri.get_bond_angles("TRI", exclude_triangle=True)) == 0
    with pytest.raises(ValueError):
        BondSet(golden / "duplicate_atoms.bnd", options)


def test_mapping_parsing(golden, options):
    # reference tests/test_mapping.py:47-87,127-130
    from pycgtool_b200 import Mapping, VirtualMap
    mapping = Mapping(golden / "water.map", options)
    assert len(mapping) == 2 and "SOL" in mapping and "HOH" in mapping
    assert mapping["SOL"][0].atoms == ["OW", "HW1", "HW2"] and mapping["HOH"][0].atoms == ["O", "H1", "H2"]
    naph = Mapping(golden / "martini3" / "naphthalene.map", options)
    assert len(naph["NAPH"]) == 5 and isinstance(naph["NAPH"][2], VirtualMap)
    assert naph["NAPH"][2].atoms == ["R1", "R2", "R4", "R5"]
    dppc = Mapping(golden / "dppc.map", options)
    assert dppc["DPPC"][0].charge == 1 and dppc["DPPC"][1].charge == -1


def test_cfg_parser_errors(golden):
    from pycgtool_b200.parsers import CFG, DuplicateSectionError, NoSectionError
    with pytest.raises(DuplicateSectionError):
        CFG(golden / "twice.cfg")
    with pytest.raises(NoSectionError):
        CFG(golden / "nosections.cfg")


def test_frame_loading(golden):
    # reference tests/test_frame.py:53-84,116-118
    from pycgtool_b200 import Frame
    frame = Frame(golden / "water.gro", golden / "water.xtc")
    assert frame.natoms == 663 and frame.n_frames == 11
    np.testing.assert_allclose(frame.atom(0).coords[0], [1.176, 1.152, 1.586], atol=1e-6)
    np.testing.assert_allclose(frame.unitcell_lengths[0], [1.9052, 1.9052, 1.9052], atol=1e-4)
    assert frame.residue(0).name == "HOH" and [a.name for a in frame.residue(0).atoms] == ["O", "H1", "H2"]
    assert Frame(golden / "polyethene.gro").unitcell_lengths is None
    from pycgtool_b200.frame import NonMatchingSystemError
    with pytest.raises(NonMatchingSystemError):
        Frame(golden / "sugar.gro", golden / "water.xtc")


def test_product_readers_agree_with_oracle_readers(golden):
    from oracle import fileio
    from pycgtool_b200 import formats
    for name in ("sugar.xtc", "water.xtc", "sugar_out.xtc"):
        mine = formats.read_xtc(golden / name)
        xyz, time, box = fileio.read_xtc(golden / name)
        assert np.array_equal(mine.xyz, xyz) and np.array_equal(mine.time, time)
        assert np.array_equal(mine.unitcell_vectors, box)


def test_synthetic_shapes():
    from workloads import synth
    assert synth.s2().n_atoms == 156672
    s5 = synth.s5(scale=0.01)
    assert s5.topology.n_residues == 4 * 92 + 13049
    s3 = synth.small_molecule()
    assert s3.n_atoms == 96 and len(s3.bnd_text.strip().splitlines()) == 1 + 23 + 22 + 21
    a = s3.coordinates_host(3, first_frame=5)
    b = s3.coordinates_host(8)
    assert np.array_equal(a, b[5:8])          # frame shards generate the same coordinates


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no product module may reference it."""
    for path in (ROOT / "pycgtool_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path


# ------------------------------------------------------------------------------------------------
# callers of the hot path: command line surface and force-field writer (no GPU involved)
# ------------------------------------------------------------------------------------------------
def test_cli_argument_surface_matches_reference():
    """reference tests/test_pycgtool.py:56-68 (test_parse_arguments) plus the defaults of
    pycgtool/__main__.py:149-254 that the hot path reads through the options object."""
    from pycgtool_b200.__main__ import parse_arguments
    args = parse_arguments(["TOPOLOGY", "-m", "MAP", "--begin", "1000"])
    assert args.topology == "TOPOLOGY" and args.trajectory is None
    assert args.mapping == "MAP" and args.bondset is None and args.begin == 1000 and args.end is None
    assert (args.map_center, args.virtual_map_center) == ("geom", "geom")
    assert args.constr_threshold == 100000 and args.temperature == 310 and args.dump_n_values == 10000
    assert args.generate_angles is True and args.generate_dihedrals is False and args.default_fc is False
    assert (args.length_form, args.angle_form, args.dihedral_form) == ("Harmonic", "CosHarmonic", "Harmonic")
    assert args.output_xtc is False and args.output_forcefield is False and args.dump_measurements is False
    flipped = parse_arguments(["T", "X", "-b", "BND", "--no-generate-angles", "--generate-dihedrals", "--output-xtc"])
    assert flipped.trajectory == "X" and flipped.generate_angles is False and flipped.generate_dihedrals is True
    assert flipped.output_xtc is True
    # a bond file without a mapping means "measure only": samples are dumped (__main__.py:257-276)
    assert flipped.dump_measurements is True
    with pytest.raises(SystemExit):
        parse_arguments(["TOPOLOGY"])          # one of -m / -b is required


def test_forcefield_rtp_r2b_from_golden_parameters(golden, options, tmp_path):
    """`.rtp` / `.r2b` text of the reference's force-field test (tests/test_pycgtool.py:175-212): the bonded
    parameters are injected from the oracle-pinned fixture, so this covers the writer alone."""
    import json
    from pycgtool_b200 import BondSet, Mapping
    from pycgtool_b200.forcefield import ForceField
    from pycgtool_b200.util import cmp_file_whitespace_float
    mapping = Mapping(golden / "sugar.map", options)
    bonds = BondSet(golden / "sugar.bnd", options)
    stored = json.loads((golden / "sugar_params.json").read_text())
    for bond, rec in zip(bonds["ALLA"], stored):
        assert list(bond.atoms) == rec["atoms"]
        bond.eqm, bond.fconst = rec["eqm"], rec["fconst"]
    ff = ForceField("out", dir_path=tmp_path)
    ff.write("out", mapping, bonds)
    out = tmp_path / "ffout.ff"
    assert cmp_file_whitespace_float(golden / "ffout.ff" / "out.rtp", out / "out.rtp", rtol=1e-6, verbose=True)
    assert (out / "out.r2b").read_text() == (golden / "ffout.ff" / "out.r2b").read_text()
    assert (out / "forcefield.itp").read_text().splitlines()[0] == "#define _FF_PYCGTOOL_out"
    # a second writer backs the directory up instead of overwriting it (util.backup_file)
    ForceField("out", dir_path=tmp_path)
    assert any(p.name.startswith("#ffout.ff") for p in tmp_path.iterdir())


def test_forcefield_terminal_residue_variants():
    """Bonds reaching into the neighbouring residue ('+' / '-' prefixes) produce N- / C- / 2-terminal building
    blocks and the matching .r2b table (pycgtool/forcefield.py:118-212)."""
    from pycgtool_b200.forcefield import ForceField

    class FakeBond:
        def __init__(self, atoms):
            self.atoms, self.eqm, self.fconst = atoms, 1.0, 100.0

        def __iter__(self):
            return iter(self.atoms)

    class FakeBonds:
        def __init__(self, table):
            self.table = table

        def get_bonds(self, mol, natoms, select=lambda x: True):
            bonds = self.table[mol]
            if natoms != -1:
                bonds = [b for b in bonds if len(b.atoms) == natoms]
            return [b for b in bonds if select(b)]

    class Bead:
        def __init__(self, name):
            self.name, self.type, self.charge = name, "P1", 0.0

    mapping = {"ETH": [Bead("C1"), Bead("C2")]}
    bonds = FakeBonds({"ETH": [FakeBond(["C1", "C2"]), FakeBond(["C2", "+C1"]), FakeBond(["-C2", "C1", "C2"])]})
    lines, nterms, cterms = ForceField.write_rtp(mapping, bonds)
    assert nterms == {"ETH"} and cterms == {"ETH"}
    text = "\n".join(lines)
    for block in ("[ ETH ]", "[ NETH ]", "[ CETH ]", "[ 2ETH ]"):
        assert block in text
    n_block = text.split("[ NETH ]")[1].split("[ CETH ]")[0]
    assert "-C2" not in n_block and "+C1" in n_block
    two_block = text.split("[ 2ETH ]")[1]
    assert "-C2" not in two_block and "+C1" not in two_block and "C1   C2" in two_block
    assert ForceField.write_r2b(nterms, cterms)[-1].split() == ["ETH", "ETH", "NETH", "CETH", "2ETH"]


def test_unique_rows_matches_numpy():
    from pycgtool_b200.mapping import _unique_rows
    rng = np.random.default_rng(0)
    pool = rng.integers(0, 50, size=(7, 12)).astype(np.int32)
    block = pool[rng.integers(0, 7, size=5000)]
    templates, inverse = _unique_rows(block)
    assert np.array_equal(templates[inverse], block)
    assert len(templates) == len(np.unique(block, axis=0))
    empty, inv = _unique_rows(np.zeros((4, 0), dtype=np.int32))
    assert len(inv) == 4 and (inv == 0).all()


def test_frame_folding_plan_for_small_systems(options):
    """Frame folding (host side): a small molecule is measured as `fold` copies per folded frame; the folded plan
    covers every (copy, column) exactly once, keeps the Bond of every column and stays one seven-group tile for the
    S3 shape; systems that already fill their warps are left alone."""
    from pycgtool_b200 import BondSet, Mapping
    from workloads import synth
    from tests.plan_decode import check_measure_plan
    sysm = synth.small_molecule()
    map_path, bnd_path = sysm.write_files()
    bondset = BondSet(bnd_path, options)
    plan = bondset.compile_plan(Mapping(map_path, options).compile_plan(sysm.topology).cg_topology)
    assert plan.fold == 9 and plan.folded is not None
    folded = plan.folded
    check_measure_plan(folded)                                   # decodes tiles / groups / lanes back to the columns
    assert folded.n_beads == 9 * plan.n_beads and folded.n_cols == 9 * plan.n_cols
    assert list(folded.tiles["ngroups"]) == [7]
    seen = np.concatenate([pos.reshape(-1) for pos in plan.fold_bond_cols])
    assert sorted(seen.tolist()) == list(range(folded.n_cols))   # every position of the folded row, once
    by_pos = np.zeros(folded.n_cols, dtype=np.int32)
    by_pos[folded.col_out] = folded.col_bond
    for bond_id, pos in enumerate(plan.fold_bond_cols):
        assert pos.shape == (9, len(plan.bond_cols[bond_id])) and (by_pos[pos] == bond_id).all()
    # frames measured through the folded plan: whole folds, and none at all for a short trajectory
    assert BondSet.folded_frames(plan, 1000) == 999 and BondSet.folded_frames(plan, 71) == 0
    bondset.fold_frames = False
    bondset._plans = {}
    assert bondset.compile_plan(Mapping(map_path, options).compile_plan(sysm.topology).cg_topology).folded is None
    # a membrane fills its warps with instances: no folding
    mem = synth.membrane(6, 0, seed=3)
    mp, bp = mem.write_files()
    assert BondSet(bp, options).compile_plan(Mapping(mp, options).compile_plan(mem.topology).cg_topology).fold == 1


def test_cg_trajectory_shares_the_unit_cell_and_bond_histogram_is_lazy():
    from pycgtool_b200.bondset import Bond
    from pycgtool_b200.functionalforms import get_functional_forms
    from pycgtool_b200.topology import Trajectory
    box = np.zeros((5, 3, 3), dtype=np.float32)
    box[:, 0, 0], box[:, 1, 1], box[:, 2, 2] = 3.0, 3.1, 3.2
    box[2, 1, 0] = 0.4
    src = Trajectory(np.zeros((5, 7, 3), dtype=np.float32), None, None, box)
    assert src.global_box_flags() == (True, True, False)
    cg = Trajectory(np.zeros((5, 2, 3), dtype=np.float32))
    assert cg.unitcell_lengths is None
    cg.share_unitcell(src)
    assert cg.unitcell_vectors is src.unitcell_vectors and cg.unitcell_lengths is src.unitcell_lengths
    assert cg.orthorhombic is False and cg.global_box_flags() == (True, True, False)
    with pytest.raises(ValueError):
        Trajectory(np.zeros((4, 2, 3), dtype=np.float32)).share_unitcell(src)
    # Bond.histogram: counts kept, edges made on demand (numpy.linspace of the range)
    bond = Bond(["A", "B"], func_form=get_functional_forms()["Harmonic"].value)
    assert bond.histogram is None
    bond._histogram = (np.arange(4, dtype=np.uint64), 0.0, 2.0)
    counts, edges = bond.histogram
    assert np.array_equal(counts, [0, 1, 2, 3]) and np.allclose(edges, [0.0, 0.5, 1.0, 1.5, 2.0])
    bond.histogram = (np.ones(2, dtype=np.uint64), np.array([1.0, 2.0, 3.0]))
    assert bond._histogram[1:] == (1.0, 3.0)

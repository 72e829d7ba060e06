"""XTC ingest: host codec vs the oracle decoder and the reference fixtures (CPU),
GPU decoder vs the host codec (gpu)."""

import numpy as np
import pytest

from oracle import fileio


def _quantised(x, precision):
    p = np.float32(precision)
    scaled = x * p
    ints = np.where(scaled >= 0, (scaled + np.float32(0.5)).astype(np.int32), (scaled - np.float32(0.5)).astype(np.int32))
    return ints.astype(np.float32) * (np.float32(1.0) / p)


@pytest.mark.parametrize("name", ["sugar.xtc", "water.xtc", "sugar_out.xtc"])
def test_host_decoder_matches_oracle_on_reference_fixtures(golden, name):
    from pycgtool_b200 import formats
    mine = formats.read_xtc(golden / name)
    xyz, time, box = fileio.read_xtc(golden / name)
    assert np.array_equal(mine.xyz, xyz) and np.array_equal(mine.time, time)
    assert np.array_equal(mine.unitcell_vectors, box)


def test_reference_kat_first_atom_of_water_xtc(golden):
    # reference tests/test_frame.py:53-56
    from pycgtool_b200 import formats
    np.testing.assert_allclose(formats.read_xtc(golden / "water.xtc").xyz[0, 0], [1.176, 1.152, 1.586], atol=1e-6)


@pytest.mark.parametrize("n_lipids,n_water,frames,precision", [(4, 30, 7, 1000.0), (1, 0, 3, 500.0), (0, 50, 5, 100.0)])
def test_encode_decode_round_trip(tmp_path, n_lipids, n_water, frames, precision):
    from pycgtool_b200 import formats
    from workloads import synth
    sysm = synth.membrane(max(n_lipids, 1), n_water, seed=3) if n_lipids else synth.membrane(1, n_water, seed=4)
    x, box = sysm.coordinates_host(frames), sysm.box_vectors(frames)
    image = formats.encode_xtc(x, box, precision=precision)
    path = tmp_path / "rt.xtc"
    image.tofile(path)
    back = formats.read_xtc(path, sysm.topology)
    assert np.array_equal(back.xyz, _quantised(x, precision))
    np.testing.assert_array_equal(back.unitcell_vectors, box)
    # the independent (oracle) decoder reads our encoder's output identically
    oxyz, _, obox = fileio.read_xtc(path)
    assert np.array_equal(oxyz, back.xyz) and np.array_equal(obox, box)
    assert len(image) < 0.6 * x.nbytes          # it actually compresses (run-length coded small differences)


def test_small_systems_are_stored_raw(tmp_path):
    from pycgtool_b200 import formats
    x = np.random.default_rng(0).normal(size=(4, 6, 3)).astype(np.float32)
    image = formats.encode_xtc(x)
    assert len(image) == 4 * (56 + 72)           # <= 9 atoms: uncompressed floats, like sugar_out.xtc
    path = tmp_path / "raw.xtc"
    image.tofile(path)
    assert np.array_equal(formats.read_xtc(path).xyz, x)


def test_wide_coordinate_range_uses_per_component_bit_fields(tmp_path):
    from pycgtool_b200 import formats
    rng = np.random.default_rng(1)
    x = rng.uniform(-9000, 9000, size=(2, 40, 3)).astype(np.float32)     # > 2^24 ints at precision 1000
    path = tmp_path / "wide.xtc"
    formats.encode_xtc(x, precision=1000.0).tofile(path)
    back = formats.read_xtc(path).xyz
    assert np.array_equal(back, fileio.read_xtc(path)[0])
    np.testing.assert_allclose(back, x, atol=2e-3)


def test_malformed_image_is_rejected(golden, tmp_path):
    from pycgtool_b200 import formats
    data = (golden / "water.xtc").read_bytes()
    bad = tmp_path / "bad.xtc"
    bad.write_bytes(data[:1000])
    with pytest.raises(OSError):
        formats.read_xtc(bad)
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(OSError):
        formats.read_xtc(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sugar.xtc", "water.xtc", "sugar_out.xtc"])
def test_gpu_decoder_matches_host_on_reference_fixtures(golden, name):
    from pycgtool_b200 import formats
    dev = formats.read_xtc_device(golden / name)
    host = formats.read_xtc(golden / name)
    assert dev.on_device
    assert np.array_equal(dev.xyz, host.xyz)
    assert np.array_equal(dev.unitcell_vectors, host.unitcell_vectors)


@pytest.mark.gpu
def test_gpu_decoder_on_synthetic_membrane_and_pipeline(tmp_path, options):
    """Compressed file -> GPU decode -> mapping equals host decode -> mapping, bit for bit."""
    from pycgtool_b200 import Frame, Mapping, formats
    from workloads import synth
    sysm = synth.membrane(6, 200, seed=8)
    frames = 37
    x, box = sysm.coordinates_host(frames), sysm.box_vectors(frames)
    path = tmp_path / "traj.xtc"
    formats.encode_xtc(x, box, precision=1000.0).tofile(path)
    dev = formats.read_xtc_device(path, sysm.topology)
    host = formats.read_xtc(path, sysm.topology)
    assert np.array_equal(dev.xyz, host.xyz)
    map_path, _ = sysm.write_files()
    mapper = Mapping(map_path, options)
    cg_dev = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=dev.device_xyz()))
    cg_host = mapper.apply(Frame.from_arrays(sysm.topology, host.xyz, unitcell_vectors=box))
    assert np.array_equal(cg_dev._trajectory.xyz, cg_host._trajectory.xyz)


@pytest.mark.gpu
def test_pipelined_upload_in_many_blocks(tmp_path):
    """Blocks of a few frames (upload on the copy stream, decode on the compute stream) give the same floats,
    and the per-block hook sees every frame exactly once, in order."""
    import torch
    from pycgtool_b200 import formats
    from workloads import synth
    sysm = synth.membrane(4, 60, seed=3)
    frames = 53
    image = formats.encode_xtc(sysm.coordinates_host(frames), sysm.box_vectors(frames), precision=1000.0)
    table, _, _, nf, na = formats._scan_xtc(image)
    host = np.empty((nf, na, 3), dtype=np.float32)
    padded = np.concatenate([image, np.zeros(8, dtype=np.uint8)])
    from pycgtool_b200 import _lib
    _lib.check(_lib.load().cgb_xtc_decode_host(_lib.ptr(padded), _lib.ptr(table), nf, na, _lib.ptr(host)))
    pinned = torch.from_numpy(padded).pin_memory()
    xyz = torch.full((nf, na, 3), float("nan"), dtype=torch.float32, device="cuda")
    seen = []
    formats.decode_image_device(pinned, image.size, table, nf, na, xyz, block_bytes=3 * (image.size // nf),
                                on_block=lambda f0, f1: seen.append((f0, f1)))
    torch.cuda.synchronize()
    assert len(seen) > 5 and seen[0][0] == 0 and seen[-1][1] == nf
    assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    assert np.array_equal(xyz.cpu().numpy(), host)


@pytest.mark.gpu
def test_frame_device_decode_flag(golden, options):
    from pycgtool_b200 import Frame, Mapping
    frame = Frame(golden / "sugar.gro", golden / "sugar.xtc", device_decode=True)
    assert frame._trajectory.on_device and frame.n_frames == 1001
    cg = Mapping(golden / "sugar.map", options).apply(frame)
    gold, _, _ = fileio.read_xtc(golden / "sugar_out.xtc")
    assert np.array_equal(cg._trajectory.xyz, gold)


def _corrupted_images(golden):
    """Bit-flipped and header-damaged variants of a reference file (deterministic)."""
    data = np.frombuffer((golden / "water.xtc").read_bytes(), dtype=np.uint8)
    rng = np.random.default_rng(2024)
    out = []
    for _ in range(60):
        img = data.copy()
        for pos in rng.integers(92, len(img), size=rng.integers(1, 40)):      # payload and later headers
            img[pos] ^= np.uint8(1 << rng.integers(0, 8))
        out.append(img)
    for field, value in ((56, 0.0), (56, -1000.0), (56, float("nan")), (56, 1e-42)):   # precision of the first frame
        img = data.copy()
        img[field:field + 4] = np.frombuffer(np.array(value, dtype=">f4").tobytes(), dtype=np.uint8)
        out.append(img)
    img = data.copy()                                                          # maxint < minint
    img[72:76] = np.frombuffer(np.array(-(2 ** 31), dtype=">i4").tobytes(), dtype=np.uint8)
    out.append(img)
    img = data.copy()                                                          # payload size says less than there is
    img[88:92] = np.frombuffer(np.array(16, dtype=">i4").tobytes(), dtype=np.uint8)
    out.append(img)
    return out


def test_host_decoder_survives_corrupt_streams(golden, tmp_path):
    """the decoder must not trust the bit stream -- a damaged file either decodes (the damage hit
    coordinate bits only) or raises; it never divides by zero, walks off the table or reads past the payload."""
    from pycgtool_b200 import formats
    path = tmp_path / "bad.xtc"
    raised = 0
    for img in _corrupted_images(golden):
        img.tofile(path)
        try:
            traj = formats.read_xtc(path)
            assert traj.xyz.shape[1:] == (traj.xyz.shape[1], 3)
        except OSError:
            raised += 1
    assert raised >= 5            # the header-damaged ones at the very least


@pytest.mark.gpu
def test_gpu_decoder_survives_corrupt_streams(golden, tmp_path):
    """The same damaged files through the GPU decoder: an OSError or a decode, and the CUDA context stays usable."""
    import torch
    from pycgtool_b200 import formats
    path = tmp_path / "bad.xtc"
    for img in _corrupted_images(golden):
        img.tofile(path)
        try:
            host = formats.read_xtc(path)
        except OSError:
            host = None
        try:
            devt = formats.read_xtc_device(path)
            torch.cuda.synchronize()
            if host is not None:
                assert np.array_equal(devt.xyz, host.xyz, equal_nan=True)
        except OSError:
            assert host is None or True
    torch.cuda.synchronize()
    good = formats.read_xtc_device(golden / "water.xtc")
    assert np.array_equal(good.xyz, formats.read_xtc(golden / "water.xtc").xyz)

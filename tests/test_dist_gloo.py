"""World-size-2 test of the frame-sharded path over gloo (CPU).

The GPU kernels are not run here; what is covered is the host logic of the
multi-GPU path: contiguous frame shards, ONE all-reduce(sum) over the packed float64
buffer [per-bond partial moments | histogram counts] (``dist.allreduce_packed``: gloo
for host tensors, the library's ``cgb_allreduce_partials`` for device tensors), the
one-pass circular variance of dihedrals, and that the merged partials invert to the
same parameters as a single pass.  The per-shard partials are produced by a
float64 NumPy statement of the accumulator definitions in
``include/pycgtool_b200.h`` (test scaffolding), the expected parameters by the oracle.
"""

import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import stats as ostats


def _partials_of(values, natoms, shift):
    """One CGB_NPART row from a value array, by definition (header, stage 2)."""
    v = values[np.isfinite(values)].astype(np.float64)
    d = v - shift
    if natoms == 4:
        d -= 2 * np.pi * np.rint(d / (2 * np.pi))       # dihedrals: wrapped about the shift
    return np.array([len(v), d.sum(), (d * d).sum(), np.cos(v).sum(), np.sin(v).sum(), 0.0, len(values), 0.0])


def _invert_row(row, natoms, shift, form, temp=310.0):
    """Host restatement of cgb_boltzmann for one row (test scaffolding only)."""
    n = row[0]
    if natoms == 2:
        md = row[1] / n
        mean, var = shift + md, row[2] / n - md * md
    else:
        mean = math.atan2(row[4] / n, row[3] / n)
        if natoms == 3:
            e = mean - shift
            var = (row[2] - 2 * e * row[1]) / n + e * e
        else:
            # one pass: exact when no sample lies between the antipodes of shift and mean (cgb_boltzmann)
            e = mean - shift
            e -= 2 * math.pi * round(e / (2 * math.pi))
            var = (row[2] - 2 * e * row[1]) / n + e * e
    eqm = ((mean) % (2 * math.pi)) - math.pi if natoms == 4 else mean
    return eqm, ostats.force_constant(form, mean, var, temp)


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pycgtool_b200 import dist
        assert dist.is_distributed() and dist.world() == (rank, world)
        rng = np.random.default_rng(123)           # same data on both ranks, each takes its shard
        n_frames, n_inst = 1001, 3
        data = {
            2: rng.normal(0.25, 0.02, size=(n_frames, n_inst)).astype(np.float32),
            3: rng.normal(1.9, 0.15, size=(n_frames, n_inst)).astype(np.float32),
            4: ((rng.normal(3.0, 0.5, size=(n_frames, n_inst)) + np.pi) % (2 * np.pi) - np.pi).astype(np.float32),
        }
        begin, end = dist.frame_shard(n_frames)
        assert (begin, end) == ((0, 501) if rank == 0 else (501, 1001))
        # shift: first value of rank 0, broadcast (BondSet.apply does the same)
        shift = torch.tensor([float(data[k][0, 0]) if rank == 0 else 0.0 for k in (2, 3, 4)], dtype=torch.float32)
        dist.broadcast(shift, src=0)
        assert shift[0].item() == float(data[2][0, 0])
        rows = [_partials_of(data[k][begin:end].reshape(-1), k, float(shift[i])) for i, k in enumerate((2, 3, 4))]
        hist = np.stack([np.histogram(data[k][begin:end], bins=16, range=r)[0] for k, r in
                         ((2, (0, 2)), (3, (0, np.pi)), (4, (-np.pi, np.pi)))]).astype(np.float64)
        # ONE packed float64 buffer [partials | hist] crosses the ranks, as in _MeasureState.invert
        packed = torch.from_numpy(np.concatenate([np.stack(rows).reshape(-1), hist.reshape(-1)]))
        dist.allreduce_packed(packed)
        partials = packed[:3 * 8].view(3, 8)
        hist = packed[3 * 8:].view(3, 16)
        assert dist.all_true(True) and not dist.all_true(rank == 0)
        # several decisions in one collective: element-wise AND over the ranks
        assert dist.all_true_many([True, rank == 0, rank == 1, True]) == (True, False, False, True)
        # the box decisions of a sharded trajectory are taken over ALL shards (mapping.py:406-408, mdtraj's
        # orthorhombic test): rank 1's shard has a zero box length in one frame and a skewed cell
        from pycgtool_b200.topology import Trajectory
        box = np.zeros((4, 3, 3), dtype=np.float32)
        box[:, 0, 0] = box[:, 1, 1] = box[:, 2, 2] = 3.0
        if rank == 1:
            box[2, 1, 0] = 0.7
        traj = Trajectory(np.zeros((4, 5, 3), dtype=np.float32), None, None, box)
        assert traj.global_box_flags() == (True, True, False)       # one rank is not orthorhombic: nobody is
        assert traj.global_box_flags() == (True, True, False)       # cached: no second collective needed
        for i, (k, form) in enumerate(((2, "Harmonic"), (3, "CosHarmonic"), (4, "Harmonic"))):
            eqm, fc = _invert_row(partials[i].numpy(), k, float(shift[i]), form)
            ref_eqm, ref_fc = ostats.boltzmann_invert(data[k].reshape(-1), k, form)
            assert abs(eqm - ref_eqm) <= 1e-4 * abs(ref_eqm) + 2e-6, (k, eqm, ref_eqm)
            assert abs(fc - ref_fc) <= 1e-4 * abs(ref_fc), (k, fc, ref_fc)
            assert int(hist[i].sum()) == int(partials[i, 0].item())
        (open(os.path.join(tmpdir, f"ok{rank}"), "w")).close()
    finally:
        td.destroy_process_group()


def test_two_rank_frame_sharding_over_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_frame_shard_partition():
    from pycgtool_b200 import dist
    for n, w in ((10, 3), (512, 8), (7, 8), (0, 2), (1001, 2)):
        blocks = [dist.frame_shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1

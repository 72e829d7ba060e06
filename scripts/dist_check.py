"""Multi-GPU check of the frame-sharded path (run under torchrun, one rank per GPU).

Every rank maps and measures its contiguous frame block of the same synthetic trajectory; the per-bond
partial moments / histograms are combined with NCCL all-reduces inside ``BondSet.boltzmann_invert``.
Rank 0 then repeats the whole trajectory alone and compares.  Prints one JSON line.
"""
import json
import os
import pathlib
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pycgtool_b200 import BondSet, Frame, Mapping, dist  # noqa: E402
from workloads import synth  # noqa: E402
from workloads.synth import Options  # noqa: E402


def main():
    rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    which = sys.argv[1] if len(sys.argv) > 1 else "s3"
    if which == "s3":
        sysm, n_frames = synth.small_molecule(), 40000
    else:
        sysm, n_frames = synth.s5(scale=0.02), 48
    opts = Options()
    opts.generate_dihedrals = True
    map_path, bnd_path = sysm.write_files()
    begin, end = dist.frame_shard(n_frames)
    xyz = sysm.coordinates_device(end - begin, dev, first_frame=begin)
    box = sysm.box_vectors(end - begin, first_frame=begin)
    mapper = Mapping(map_path, opts)
    bonds = BondSet(bnd_path, opts)
    td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cg = mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box, xyz_device=xyz))
    bonds.apply(cg)
    bonds.boltzmann_invert()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sharded = [(b.eqm, b.fconst) for mol in bonds for b in bonds[mol]]
    hist_total = int(bonds._state.hist.sum().item())
    result = None
    if rank == 0:
        # single-GPU run over the whole trajectory; torch.distributed must not be consulted here
        saved = dist.is_distributed
        dist.is_distributed = lambda: False
        try:
            xyz_all = sysm.coordinates_device(n_frames, dev)
            box_all = sysm.box_vectors(n_frames)
            whole = BondSet(bnd_path, opts)
            whole.apply(mapper.apply(Frame.from_arrays(sysm.topology, unitcell_vectors=box_all, xyz_device=xyz_all)))
            whole.boltzmann_invert()
        finally:
            dist.is_distributed = saved
        single = [(b.eqm, b.fconst) for mol in whole for b in whole[mol]]
        worst = 0.0
        for (e1, f1), (e2, f2) in zip(sharded, single):
            worst = max(worst, abs(e1 - e2) / max(abs(e2), 1e-9), abs(f1 - f2) / max(abs(f2), 1e-9))
        result = {"check": "sharded == single", "world": world, "workload": which, "frames": n_frames,
                  "bonds": len(single), "max_rel_diff": worst, "hist_total": hist_total,
                  "hist_expected": int(whole._state.hist.sum().item()), "ok": bool(worst < 1e-6),
                  "pipeline_s": dt}
        print(json.dumps(result), flush=True)
    td.barrier()
    td.destroy_process_group()
    if result is not None and not (result["ok"] and result["hist_total"] == result["hist_expected"]):
        sys.exit(1)


if __name__ == "__main__":
    main()

"""Slab-decomposed inverse Laplacian over the ranks of a torchrun job (one process per GPU):
correctness against the single-GPU solve and time per solve (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/slab_bench.py --grid 4096 --reps 10
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import _lib as L  # noqa: E402
from ilm_b200 import shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--in-library", action="store_true", help="exchanges by ilm_slab_solve (grouped ncclSend/ncclRecv inside the library)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", rank=rank, world_size=world)
    g = ilm.PhysicalGrid.centered(a.grid)
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g, device=True)
    slab = shard.SlabLaplacian(cache, L.NODES_PRIMAL)
    if a.in_library:
        cache.comm_init()
    # a slab-only rank does not need the full-size spectrum buffers of the plan: release them and look at the footprint
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    L.check(cache._lib.ilm_plan_release_spectrum(cache._plan))
    freed_gb = (torch.cuda.mem_get_info()[0] - free0) / 1e9
    shape = g.layout_shape(L.NODES_PRIMAL)
    w = np.random.default_rng(0).standard_normal(shape)
    mine = slab.scatter(w)
    keep = mine.clone()
    slab.inverse_laplacian(mine, in_library=a.in_library)
    torch.cuda.synchronize()
    slab_only_gb = (free0 + int(freed_gb * 1e9) - torch.cuda.mem_get_info()[0]) / 1e9     # what the slab solve allocated
    full = ilm.Nodes(ilm.Primal, g, device=True).set(w)
    ilm.inverse_laplacian(full, cache)
    r0, r1 = slab.rows(L.NODES_PRIMAL)
    ref = torch.from_numpy(np.ascontiguousarray(full.array()[:, r0:r1].T).reshape(-1)).to(mine.device)
    same = bool(torch.equal(mine, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        mine.copy_(keep)
        slab.inverse_laplacian(mine, in_library=a.in_library)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        slab.inverse_laplacian(mine, in_library=a.in_library)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device=mine.device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # single-GPU time of the same solve on this rank, for the scaling ratio
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        ilm.inverse_laplacian(full, cache)
    e1.record()
    torch.cuda.synchronize()
    single = e0.elapsed_time(e1) / a.reps
    ok = torch.tensor([1.0 if same else 0.0], device=mine.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "slab-decomposed inverse Laplacian, one real field", "grid": a.grid, "n_gpus": world,
                          "ms_per_solve_slab": float(ms.item()), "ms_per_solve_single_gpu": single,
                          "bit_identical_to_single_gpu": bool(ok.item() == 1.0), "exchange": "in-library NCCL" if a.in_library else "torch all_to_all_single",
                          "full_spectrum_buffers_released_GB": freed_gb, "slab_buffers_allocated_GB": slab_only_gb,
                          "exchange_bytes_per_rank_per_solve": 2 * 8 * sum(slab.counts[0][0])}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

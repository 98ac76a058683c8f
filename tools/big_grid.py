"""Largest single grid (BASELINE config C5b): one LGF solve on n x n (n = 8192 / 16384), per-pass
times from ilm_profile_conv (CUDA events on the plan's stream) and, under torchrun, the
slab-decomposed solve over the ranks.

    python tools/big_grid.py --grid 16384 --reps 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/big_grid.py --grid 16384 --reps 5
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import _lib as L  # noqa: E402
from ilm_b200 import shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world)
    n = a.grid
    t0 = time.time()
    G = ilm.lgf.lgf_table(n)
    t_table = time.time() - t0
    g = ilm.PhysicalGrid.centered(n)
    t0 = time.time()
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g, lgf_table=G, device=True)
    cache.sync()
    t_plan = time.time() - t0
    del G
    P = n * n
    ms = (C.c_double * 3)()
    L.check(cache._lib.ilm_profile_conv(cache._plan, L.NODES_PRIMAL, a.reps, C.byref(ms)))
    Lh = 1 << int(np.ceil(np.log2(n)))              # half padded length
    spec = 2 * Lh * n * 16                          # one spectrum buffer, bytes
    ghat = (Lh + 1) * 2 * Lh * 8
    bytes_pass = [2 * 8 * P + spec, spec + ghat + spec, spec + 2 * 8 * P]
    out = {"what": "LGF inverse Laplacian, two real fields per complex transform", "grid": n, "n_gpus": world,
           "half_padded_length": Lh, "lgf_table_s": t_table, "plan_create_s": t_plan,
           "passes_ms": {"A_rows_fwd": ms[0], "B_columns": ms[1], "C_rows_inv": ms[2]},
           "passes_GBps": {k: b / (m * 1e-3) / 1e9 for k, b, m in zip("ABC", bytes_pass, ms)},
           "pair_solve_ms": sum(ms), "grid_point_solves_per_s": 2 * P / (sum(ms) * 1e-3),
           "algorithmic_bytes_per_pair_solve": sum(bytes_pass),
           "frac_of_hbm_peak_6540": sum(bytes_pass) / (sum(ms) * 1e-3) / 1e9 / 6540.8}
    # correctness at full size: L (L^-1 w) = w away from the boundary
    w = torch.zeros(n - 1, n - 1, dtype=torch.float64)
    c = n // 2
    w[c - 100:c + 100, c - 100:c + 100] = torch.randn(200, 200, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    d = ilm.Nodes(ilm.Primal, g, device=True).set(w.numpy().T)
    ilm.inverse_laplacian(d, cache)
    lap = ilm.Nodes(ilm.Primal, g, device=True)
    ilm.laplacian(lap, d, cache)
    res = (lap.array() - w.numpy().T)[1:-1, 1:-1]
    out["max_residual_L_Linv_w"] = float(np.abs(res).max())
    if world > 1:
        import torch.distributed as dist
        slab = shard.SlabLaplacian(cache, L.NODES_PRIMAL)
        mine = slab.scatter(w.numpy().T)
        slab.inverse_laplacian(mine)
        r0, r1 = slab.rows(L.NODES_PRIMAL)
        ref = torch.from_numpy(np.ascontiguousarray(d.array()[:, r0:r1].T).reshape(-1)).to(mine.device)
        ok = torch.tensor([1.0 if torch.equal(mine, ref) else 0.0], device=mine.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        slab.inverse_laplacian(mine)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            slab.inverse_laplacian(mine)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.reps], device=mine.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # single-GPU time of one real field through the public call, for the ratio
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            ilm.inverse_laplacian(d, cache)
        e1.record()
        torch.cuda.synchronize()
        out["slab"] = {"ms_per_solve": float(t.item()), "single_gpu_ms_per_solve": e0.elapsed_time(e1) / a.reps,
                       "bit_identical_to_single_gpu": bool(ok.item() == 1.0),
                       "exchange_bytes_per_rank_per_solve": 2 * 8 * sum(slab.counts[0][0])}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""BASELINE config C5a: an ensemble of independent Stokes / vector-face (Rf) problems on 1024^2 grids,
sharded over the ranks as replicas (problem index modulo world size, no data-path collective; one
all_reduce of the timing at the end).  Each problem: SurfaceVectorCache of a rectangle at its own
position, S = create_CL2invCT (2N x 2N, two inverse Laplacians per column), Ss = create_CLinvCT_scalar,
their LU factorisations and the Stokes solve of test/literate/stokes.jl:98-166.

    python tools/ensemble_c5a.py --problems 8
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tools/ensemble_c5a.py --problems 512
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilm_b200 as ilm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1024)
ap.add_argument("--problems", type=int, default=8)
a = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world)
g = ilm.PhysicalGrid.centered(a.grid)
G = ilm.lgf.lgf_table(a.grid)
rng = np.random.default_rng(0)
centers = rng.uniform(-1.0, 1.0, size=(a.problems, 2))          # same list on every rank
mine = [k for k in range(a.problems) if k % world == rank]


def one(k, cache=None):
    body = ilm.bodies.rectangle(0.5, 0.25, 1.4 * g.dx, center=tuple(centers[k]))
    if cache is None:
        cache = ilm.SurfaceVectorCache(body, g, lgf_table=G, device=True)
    else:
        cache.update_points(body)                                # same L (Ghat), new tables
    N = cache.N
    v, s, sigma, S, Ss = ilm.stokes_flow(cache, np.concatenate([np.ones(N), np.zeros(N)]))
    return cache, float(sigma.data[:N].sum().item()) * 0.0 + float(v.data.abs().max().item()), N


cache, _, N = one(mine[0] if mine else 0)                        # warm-up (plan creation, first launches)
torch.cuda.synchronize()
l0 = cache.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
vmax = 0.0
for k in mine:
    _, m, N = one(k, cache)
    vmax = max(vmax, m)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
solves_per_problem = 2 * (2 * N) + N + 7                          # CL2invCT columns (2 each), CLinvCT_scalar, the solve itself
if rank == 0:
    t = float(ms.item()) * 1e-3
    print(json.dumps({"config": "C5a ensemble of independent Stokes problems (replicas, no collective)", "grid": a.grid,
                      "problems": a.problems, "n_gpus": world, "surface_points": N, "seconds": t,
                      "problems_per_s": a.problems / t, "inverse_laplacians_per_problem": solves_per_problem,
                      "grid_point_solves_per_s": a.problems * solves_per_problem * a.grid ** 2 / t,
                      "launches_per_problem": (cache.launch_count() - l0) / max(len(mine), 1), "max_abs_velocity": vmax}))
if world > 1:
    dist.destroy_process_group()

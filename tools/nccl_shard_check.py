"""Multi-GPU check of the in-library NCCL path (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/nccl_shard_check.py [--grid 1024]

* ilm_create_schur_sharded over the plan's communicator is BIT-EQUAL to the single-rank ilm_create_schur (same kernels
  on the same columns; the exchange is a grouped in-place ncclBroadcast), and equal to the torch all-gather path;
* ilm_dirichlet_poisson with a communicator returns the single-rank field / multiplier on every rank;
* ilm_slab_solve (grouped ncclSend / ncclRecv inside the library) is bit-equal to the single-GPU inverse Laplacian.
Prints one JSON line from rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import _lib as L, shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    g = ilm.PhysicalGrid.centered(args.grid)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(args.grid, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
    N = cache.N
    fplus = cache.points()[0].copy()
    # single-rank results first (no communicator yet)
    S1 = ilm.create_RTLinvR(cache).clone()
    f1, s1 = ilm.dirichlet_solve(cache, fplus)
    f1, s1 = f1.data.clone(), s1.data.clone()
    cache.comm_init()
    assert cache.comm_info() == (rank, world)
    S2 = ilm.create_schur_sharded(cache, "RTLinvR")
    S3 = shard.create_schur_sharded(ilm.create_RTLinvR, cache)
    f2, s2 = ilm.dirichlet_solve(cache, fplus)
    out = {"world": world, "grid": args.grid, "N": N,
           "schur_sharded_bit_equal": bool(torch.equal(S1, S2)),
           # the torch path builds column ranges (every column over all window rows), the library builds the whole
           # matrix with the symmetric half-row probes: equal to rounding, not bit for bit
           "schur_torch_path_rel_diff": float((S1 - S3).abs().max() / S1.abs().max()),
           "schur_torch_path_bit_equal": bool((S1 - S3).abs().max() <= 1e-13 * S1.abs().max()),
           "dirichlet_f_bit_equal": bool(torch.equal(f1, f2.data)), "dirichlet_s_bit_equal": bool(torch.equal(s1, s2.data))}
    # slab solve inside the library vs the single-GPU solve
    rng = np.random.default_rng(3)
    w = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    full = cache.zeros_grid().set(w)
    ilm.inverse_laplacian(full, cache)
    slab = shard.SlabLaplacian(cache, L.NODES_PRIMAL)
    mine = slab.scatter(w)
    mine = slab.inverse_laplacian(mine, in_library=True)
    ref = slab.scatter(full.array())
    out["slab_in_library_bit_equal"] = bool(torch.equal(mine, ref))
    flags = torch.tensor([int(v) for k, v in out.items() if k.endswith("bit_equal")], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(flags.min().item() == 1)
    cache.comm_destroy()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()

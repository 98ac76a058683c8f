"""Wall-clock pieces of one bench step per rank (diagnostic; under torchrun for N > 1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ilm_b200 as ilm

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = ilm.PhysicalGrid.centered(n)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(n, cache_dir="/tmp/ilm_lgf_cache")
cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
if world > 1:
    cache.comm_init()
fplus = torch.from_numpy(cache.points()[0].copy()).cuda()
for it in range(5):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    cache.update_points(body)
    t1 = time.perf_counter()
    f, s = ilm.dirichlet_solve(cache, fplus)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"rank {rank} step {it}: update_points {1e3 * (t1 - t0):.2f} ms (returns after its sync), dirichlet_solve enqueue "
          f"{1e3 * (t2 - t1):.2f} ms, drain {1e3 * (t3 - t2):.2f} ms, total {1e3 * (t3 - t0):.2f} ms", flush=True)

# back to back, as bench.py times it (CUDA events around K steps, no barrier between steps)
for K in (3, 3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        cache.update_points(body)
        f, s = ilm.dirichlet_solve(cache, fplus)
    e1.record()
    torch.cuda.synchronize()
    print(f"rank {rank}: {K} steps back to back: {e0.elapsed_time(e1) / K:.2f} ms per step (events), {1e3 * (time.perf_counter() - t0) / K:.2f} ms (wall)", flush=True)

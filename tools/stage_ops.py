"""One launch of every non-FFT stage of the path at BASELINE size (4096^2, circle, N = 4593) for ncu captures
(dram__bytes_read/write.sum per kernel next to the algorithmic bytes of bench.py's `stages`), plus the LU kernels
(FP64 / tensor pipe counters).  Numbers printed under a profiler are never bench values.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/stage_traffic.csv python tools/stage_ops.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ilm_b200 as ilm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=4096)
ap.add_argument("--lu", action="store_true", help="also factor and solve an N x N Schur complement (direct-table build)")
args = ap.parse_args()
g = ilm.PhysicalGrid.centered(args.grid)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(args.grid, cache_dir="/tmp/ilm_lgf_cache")
cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
gen = torch.Generator(device="cuda").manual_seed(0)
qe, pn, sd = cache.zeros_gridgrad(), cache.zeros_grid(), cache.zeros_gridcurl()
for t in (qe, pn, sd):
    t.data.normal_(generator=gen)
oq, op_, os_ = cache.zeros_gridgrad(), cache.zeros_grid(), cache.zeros_gridcurl()
fs, fo = cache.zeros_surface(), cache.zeros_surface()
fs.data.normal_(generator=gen)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
ops = [lambda: ilm.divergence(op_, qe, cache), lambda: ilm.grad(oq, pn, cache), lambda: ilm.curl(oq, sd, cache),
       lambda: ilm.curl(os_, qe, cache), lambda: ilm.laplacian(os_, sd, cache), lambda: ilm.regularize(op_, fs, cache),
       lambda: ilm.interpolate(fo, pn, cache), lambda: ilm.surface_divergence(op_, fs, cache),
       lambda: ilm.surface_grad(fo, pn, cache)]
for fn in ops:
    flush.zero_()
    fn()
if args.lu:
    S = ilm.create_RTLinvR_direct(cache)
    lu = ilm.LU(S)
    lu.solve(torch.randn(cache.N, dtype=torch.float64, device="cuda"))
cache.sync()
print("done", cache.N)

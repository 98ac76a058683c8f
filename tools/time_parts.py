"""Time the parts of one Dirichlet step on the device (tools, not product)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ilm_b200 as ilm

NG = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = ilm.PhysicalGrid.centered(NG)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)

def timed(name, fn, reps=1):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.perf_counter()-t0)/reps*1e3:10.3f} ms")
    return out

S = timed("create_RTLinvR", lambda: ilm.create_RTLinvR(cache))
Sd = timed("create_RTLinvR_direct", lambda: ilm.create_RTLinvR_direct(cache))
print("direct vs column-solve rel diff", ((Sd - S).abs().max() / S.abs().max()).item())
lu = timed("LU factor", lambda: ilm.LU(S))
b = torch.randn(cache.N, dtype=torch.float64, device="cuda")
x = timed("LU solve", lambda: lu.solve(b), reps=3)
r = (S @ x - b).abs().max().item() / b.abs().max().item()
print("solve residual", r)
w = cache.zeros_grid(); w.data.normal_()
timed("inverse_laplacian (1 field)", lambda: ilm.inverse_laplacian(w, cache), reps=5)
f = cache.zeros_surface(); f.data.normal_()
timed("surface_divergence", lambda: ilm.surface_divergence(w, f, cache), reps=5)
timed("create_surface_filter", lambda: ilm.create_surface_filter(cache))
C = ilm.create_surface_filter(cache)
timed("matvec_pow k=5", lambda: ilm.matvec_pow(C, 5, f), reps=3)
timed("mask", lambda: ilm.mask(cache), reps=3)

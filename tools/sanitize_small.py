"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ilm_b200 as ilm
for n in (48, 300):                     # L = 64 (F > 1 path) and L = 512 (TMA path)
    g = ilm.PhysicalGrid.centered(n)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    cache = ilm.SurfaceScalarCache(body, g)
    f, s, S = ilm.dirichlet_poisson(cache, cache.points()[0].copy(), filter_passes=2)
    Sd = ilm.create_RTLinvR_direct(cache)
    print(n, cache.N, float(np.abs(S - Sd).max() / np.abs(S).max()))
    w = cache.zeros_gridgrad(); w.data[:] = 1.0
    ilm.inverse_laplacian(w, cache)
    m = ilm.mask(cache)
    A = ilm.create_GLinvD(cache, cols=(0, 6))
    vc = ilm.SurfaceVectorCache(body, g)
    B = ilm.create_GLinvD(vc, cols=(0, 4))
print("done")

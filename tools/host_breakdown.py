import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import ilm_b200 as ilm
from ilm_b200 import api
g = ilm.PhysicalGrid.centered(4096)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(4096)
hc = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=False)
N = hc.N
def T(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
out = {}
out['host_zeros_NxN'] = T(lambda: api._host_zeros(N * N))
S = ilm.create_RTLinvR(hc, cols=(0, 64))
out['create_RTLinvR_64cols'] = T(lambda: ilm.create_RTLinvR(hc, cols=(0, 64)))
Sfull = np.asfortranarray(np.random.default_rng(0).standard_normal((N, N)) + N * np.eye(N))
out['LU_host'] = T(lambda: ilm.LU(Sfull))
lu = ilm.LU(Sfull)
b = np.ones(N)
out['solve_host'] = T(lambda: lu.solve(b))
w = hc.zeros_grid(); w.data[...] = 1.0
out['inverse_laplacian_host'] = T(lambda: ilm.inverse_laplacian(w, hc))
d = hc.zeros_surface().set(np.ones(N))
out['surface_divergence_host'] = T(lambda: ilm.surface_divergence(w, d, hc))
out['zeros_grid'] = T(lambda: hc.zeros_grid())
print(json.dumps(out))

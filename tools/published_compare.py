"""The only timings the reference publishes (`@time` outputs of executed
notebooks, Julia 1.9.2, unknown CPU, grid 406 x 406 from PhysicalGrid((-2,2),(-2,2),0.01)) next to the
same calls on one B200 -- same grid, same bodies, S prebuilt as in the notebooks.  One JSON line.
Not the headline metric (bench.py); a sanity check against numbers that exist upstream."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import _lib as L  # noqa: E402

PUBLISHED_MS = {"neumann_solve": 55.7, "stokes_solve": 96.3, "convective_derivative_scalar": 3.82,
                "convective_derivative_vector": 10.7, "w_cross_v": 1.62}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dx = 0.01
    g = ilm.PhysicalGrid(406, 406, dx, (203, 203))                    # examples/caches.ipynb cell 5
    G = ilm.lgf.lgf_table(406)
    out = {"grid": [406, 406], "published_ms": PUBLISHED_MS, "published_source": "BASELINE.md section 1 (unknown CPU, Julia 1.9.2)"}
    res = {}
    for device in (True, False):
        tag = "device_resident" if device else "host_buffers"
        r = {}
        # Neumann: circle R = 1, 448 points (examples/neumann.ipynb), S = create_CLinvCT prebuilt
        sc = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * dx), g, lgf_table=G, device=device)
        lu = ilm.LU(ilm.create_CLinvCT(sc))
        vn = sc.normals()[0].copy()
        r["neumann_solve"] = timed(lambda: ilm.neumann_poisson(sc, vn, S=lu), 10)
        # Stokes: rectangle 0.5 x 0.25 (examples/stokes.ipynb), S, Ss prebuilt
        vc = ilm.SurfaceVectorCache(ilm.bodies.rectangle(0.5, 0.25, 1.4 * dx), g, lgf_table=G, device=device)
        N = vc.N
        vplus = np.concatenate([np.ones(N), np.zeros(N)])
        _, _, _, S, Ss = ilm.stokes_flow(vc, vplus)
        r["stokes_solve"] = timed(lambda: ilm.stokes_flow(vc, vplus, S=S, Ss=Ss), 10)
        # convective terms on random fields (examples/gridops.ipynb)
        q, res_e = vc.zeros_grid(), vc.zeros_grid()
        p, res_p, w = vc.zeros_griddiv(), vc.zeros_griddiv(), vc.zeros_gridcurl()
        rng = np.random.default_rng(0)
        for d in (q, p, w):
            d.set(rng.standard_normal(len(d)))
        r["convective_derivative_scalar"] = timed(lambda: ilm.convective_derivative(res_p, q, p, vc), 50)
        r["convective_derivative_vector"] = timed(lambda: ilm.convective_derivative(res_e, q, vc), 50)
        r["w_cross_v"] = timed(lambda: ilm.w_cross_v(res_e, w, q, vc), 50)
        r["surface_points"] = {"circle": sc.N, "rectangle": N}
        res[tag] = r
    out["b200_ms"] = res
    out["speedup_vs_published_host_buffers"] = {k: PUBLISHED_MS[k] / res["host_buffers"][k] for k in PUBLISHED_MS}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

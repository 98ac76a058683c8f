"""BASELINE config C3: unsteady heat conduction with the integrating factor (plan_intfact), 2048^2
grid, moving circle, IF-HERK (timemarching.DirichletHeatConduction).  Every step refreshes the plan
(tables of the moved body), rebuilds and factors the two stage complements, and applies 7
integrating-factor convolutions.  Prints one JSON line (CUDA-event time per step).

    python tools/heat_c3.py --grid 2048 --steps 1000
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import timemarching as tm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=2048)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--probe", action="store_true", help="build the stage complements column by column (reference's way)")
ap.add_argument("--breakdown", action="store_true", help="time the pieces of a step separately (20 repetitions each)")
a = ap.parse_args()
g = ilm.PhysicalGrid.centered(a.grid)
dt = tm.timestep_fourier(g, 1.0, 1.0)
# x_c(t) = -0.5 + U t; U chosen so that the body crosses one unit length over the run (x_c = -0.5 + t)
U = 1.0 / (a.steps * dt)


def body_at(t):
    return ilm.bodies.circle(1.0, 1.4 * g.dx, center=(-0.5 + U * t, 0.0))


prob = tm.DirichletHeatConduction(g, body_at, kappa=1.0, fourier=1.0, Tplus=0.0, Tminus=1.0, moving=True,
                                  direct_schur=not a.probe, device=True)
for _ in range(3):
    prob.step()
torch.cuda.synchronize()
if a.breakdown:
    import time

    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    c = prob.cache
    a_half = max(prob.stage_a)
    S = ilm.create_RTHR_direct(c, prob.table[a_half])
    w = c.zeros_grid()
    w.data.normal_()
    sd = c.zeros_surface()
    print(json.dumps({"breakdown_ms": {
        "update_points (table rebuild)": timed(lambda: c.update_points(body_at(prob.t))),
        "create_RTHR_direct (one stage complement)": timed(lambda: ilm.create_RTHR_direct(c, prob.table[a_half])),
        "LU factorisation (N = %d)" % c.N: timed(lambda: ilm.LU(S)),
        "intfact convolution (one field)": timed(lambda: ilm.convolve(w, c, prob.kernel_id[a_half])),
        "ode_rhs (surface_divergence)": timed(lambda: prob.ode_rhs(0.0)),
        "interpolate + regularize": timed(lambda: (ilm.interpolate(sd, w, c), ilm.regularize(w, sd, c))),
        "whole step": timed(prob.step),
    }, "per_step_counts": {"update_points": 1, "create_RTHR_direct": 2, "LU": 2, "intfact convolution": 7, "ode_rhs": 3,
                           "interpolate + regularize": 3}}))
    sys.exit(0)
l0 = prob.cache.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
prob.run(a.steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
T = prob.T.array()
c = g.I0[0] - 1
xc = -0.5 + U * prob.t
ic = int(round(xc / g.dx)) + c
print(json.dumps({"config": "C3 heat conduction, IF-HERK (3 stages), moving circle", "grid": a.grid, "surface_points": prob.cache.N,
                  "steps": a.steps, "ms_per_step": ms, "stage_complements": "direct table" if not a.probe else "column probes",
                  "inverse_or_intfact_convolutions_per_step": 7, "plan_refreshes": prob.stats["plan_refreshes"],
                  "schur_builds": prob.stats["schur_builds"], "launches_per_step": (prob.cache.launch_count() - l0) / a.steps,
                  "grid_point_convolutions_per_s": 7 * a.grid ** 2 / (ms * 1e-3),
                  "T_at_body_centre": float(T[ic, c]), "T_far_corner": float(T[2, 2]), "finite": bool(np.isfinite(T).all())}))

"""Drive the three FFT-convolution passes a few times on a BASELINE-size plan
(for ncu captures and launch lists; numbers printed under a profiler are never
bench values)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilm_b200 as ilm  # noqa: E402
from ilm_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=4096)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--schur-cols", type=int, default=0)
ap.add_argument("--probe", action="store_true", help="also time the passes in Schur-probe mode")
args = ap.parse_args()
g = ilm.PhysicalGrid.centered(args.grid)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(args.grid, cache_dir="/tmp/ilm_lgf_cache")
cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
ms = (L.C.c_double * 3)()
L.check(cache._lib.ilm_profile_conv(cache._plan, L.NODES_PRIMAL, args.reps, L.C.byref(ms)))
print("pass A/B/C ms per launch:", [round(x, 4) for x in ms])
if args.probe:
    mq = (L.C.c_double * 3)()
    L.check(cache._lib.ilm_profile_conv_probe(cache._plan, cache.N // 3, args.reps, L.C.byref(mq)))
    print("probe-mode pass A/B/C ms per launch:", [round(x, 4) for x in mq])
if args.schur_cols:
    S = ilm.create_RTLinvR(cache, cols=(0, args.schur_cols))
    cache.sync()
    print("schur block", tuple(S.shape))

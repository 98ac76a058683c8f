"""Device timings of the forcing and Helmholtz entry points at full size (tools, not product).
usage: python tools/forcing_helmholtz_bench.py [NG] [out.json]; CUDA events on the plans' (default) stream."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ilm_b200 as ilm

NG = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
peak = 6534.5
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = ilm.PhysicalGrid.centered(NG)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
sc = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
vc = ilm.SurfaceVectorCache(body, g, device=True, parent=sc)
P = NG * NG
res = {"grid": NG, "N": sc.N, "hbm_peak_gbs": peak}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # 256 MiB > 126 MB L2


def timed(name, fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = nbytes / ms * 1e-6
    res[name] = {"ms": ms, "algorithmic_bytes": int(nbytes), "GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    print(f"{name:34s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f} % of HBM peak")


# forcing: one area heater (circle R = 0.5), one line heater, two point sources
shape = ilm.bodies.circle(0.5, 1.4 * g.dx, center=(0.3, 0.2))
fc = ilm.ForcingModelAndRegion([ilm.AreaForcingModel(shape, ilm.RigidTransform(), lambda s, T, t, fr, pp: None),
                                ilm.LineForcingModel(shape, ilm.RigidTransform(), lambda s, T, t, fr, pp: None),
                                ilm.PointForcingModel((np.array([-1.2, 0.5]), np.array([0.5, 0.5])), lambda s, T, t, fr, pp: None,
                                                      ddftype="m4prime")], sc)
area, line, pts = (f.region_cache for f in fc)
area.str.data.normal_(); line.str.data.normal_(); pts.str.data.normal_()
dT = sc.zeros_grid()
lib, ptr = sc._lib, ilm.api._ptr
npr = len(dT)
timed("forcing_area_add (str*mask)", lambda: lib.ilm_forcing_area_add(area.cache._plan, 0, ptr(area.str.data), ptr(area.mask.data), ptr(dT.data)), 32 * npr)
timed("forcing_area_add (whole domain)", lambda: lib.ilm_forcing_area_add(area.cache._plan, 0, ptr(area.str.data), None, ptr(dT.data)), 24 * npr)
nl = line.cache.N
timed("forcing_line_add", lambda: lib.ilm_forcing_line_add(line.cache._plan, 0, ptr(line.str.data), ptr(dT.data)), nl * (16 * 12 + 8) + 16 * 16 * nl)
timed("forcing_point_add (2 points)", lambda: lib.ilm_forcing_line_add(pts.cache._plan, 0, ptr(pts.str.data), ptr(dT.data)), 2 * (16 * 12 + 8 + 16 * 16))
# the reference's composition of the line forcing: fill + regularize + whole-field add
tmp = sc.zeros_grid()
timed("reference-style line forcing", lambda: (ilm.regularize(tmp, line.str, line.cache), dT.data.add_(tmp.data)), 8 * npr + 24 * npr)

# Helmholtz
w, d, dv = vc.zeros_gridcurl(), vc.zeros_griddiv(), vc.zeros_surface()
w.data.normal_(); d.data.normal_(); dv.data.normal_()
psi, phi, v, vp = vc.zeros_gridcurl(), vc.zeros_griddiv(), vc.zeros_grid(), vc.zeros_grid()
vp.data.normal_()
timed("helmholtz potentials (pair solve)", lambda: ilm.potentials_from_masked_fields(psi, phi, w, d, dv, vc), 2 * 88 * P + 32 * P)
timed("vecfield_from_potentials", lambda: lib.ilm_vecfield_from_potentials(vc._plan, ptr(psi.data), ptr(phi.data), ptr(vp.data), ptr(v.data)), 48 * P)
timed("vecfield_helmholtz (fused)", lambda: ilm.vecfield_helmholtz(v, w, d, dv, vp, vc), 2 * 88 * P + 32 * P + 48 * P)


def reference_style():            # the reference's sequence of whole-field operations through the same library
    t1, t2 = vc.zeros_gridcurl(), vc.zeros_griddiv()
    ilm.regularize_normal_cross(t1, dv, vc); t1.data.add_(w.data); ilm.inverse_laplacian(t1, vc); t1.data.mul_(-1.0)
    ilm.regularize_normal_dot(t2, dv, vc); t2.data.add_(d.data); ilm.inverse_laplacian(t2, vc)
    a, b = vc.zeros_grid(), vc.zeros_grid()
    ilm.curl(a, t1, vc); ilm.grad(b, t2, vc)
    v.data.copy_(a.data + b.data); v.data.add_(vp.data)


timed("vecfield_helmholtz (reference-style sequence)", reference_style, 2 * 88 * P + 32 * P + 48 * P, reps=5)
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)

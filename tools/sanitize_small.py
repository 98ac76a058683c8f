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

# ---- features added later in round 1: mask products, convective terms, slab stages, big lengths, IF-HERK
import ctypes as C  # noqa: E402
from ilm_b200 import _lib as L, shard, timemarching as tm  # noqa: E402

g = ilm.PhysicalGrid.centered(48)
body = ilm.bodies.circle(1.0, 1.4 * g.dx)
cache = ilm.SurfaceScalarCache(body, g)
vc = ilm.SurfaceVectorCache(body, g)
for w in (ilm.Nodes(ilm.Primal, g), ilm.Nodes(ilm.Dual, g), ilm.XEdges(g), ilm.YEdges(g), ilm.Edges(g)):
    ilm.mask(w.fill(1.0), cache)
    ilm.complementary_mask(w, cache)
for w in (ilm.Edges(g), ilm.Nodes(ilm.Primal, g), ilm.Nodes(ilm.Dual, g), ilm.EdgeGradient(g)):
    ilm.mask(w.fill(1.0), vc)
q, r = ilm.Edges(g).fill(0.5), ilm.Edges(g)
ilm.convective_derivative(r, q, cache)
ilm.convective_derivative(ilm.Nodes(ilm.Primal, g), q, ilm.Nodes(ilm.Primal, g).fill(2.0), cache)
ilm.w_cross_v(r, ilm.Nodes(ilm.Dual, g).fill(1.0), q, cache)
v, s, sig, S, Ss = ilm.stokes_flow(vc, np.concatenate([np.ones(vc.N), np.zeros(vc.N)]))
prob = tm.DirichletHeatConduction(g, lambda t: ilm.bodies.circle(0.7, 1.4 * g.dx, center=(-0.3 + 20 * t, 0.0)), moving=True, device=True)
prob.run(2)
prob = tm.DirichletHeatConduction(g, lambda t: body, direct_schur=False, device=False)
prob.run(1)
# slab stages for 3 virtual ranks on a device-resident plan (pack / unpack copies, ranged column pass)
virtual_slab_solve = shard.slab_solve_virtual_ranks
for NX, NY in ((48, 48), (40, 600), (4200, 24), (24, 4200)):          # the last two: half length 8192 (radix-2Q bodies)
    gg = ilm.PhysicalGrid(NX, NY, 0.05, (NX // 2, NY // 2))
    dc = ilm.SurfaceScalarCache(ilm.bodies.circle(0.4, 0.07), gg, lgf_table=ilm.lgf.lgf_table(max(NX, NY)), device=True)
    wf = np.random.default_rng(0).standard_normal(gg.layout_shape(L.NODES_PRIMAL))
    virtual_slab_solve(dc, [L.NODES_PRIMAL], [wf], 3)
    ilm.create_RTLinvR(dc, cols=(0, 4))
print("done (round-1 additions)")

# ---- forcing regions and Helmholtz decomposition (shared plans, accumulate-gather, fused recomposition)
for dev in (False, True):
    gg = ilm.PhysicalGrid(61, 53, 0.07, (30, 26))
    bb = ilm.bodies.circle(1.0, 0.1)
    scc = ilm.SurfaceScalarCache(bb, gg, device=dev)
    vcc = ilm.SurfaceVectorCache(bb, gg, device=dev, parent=scc)
    shp = ilm.bodies.rectangle(0.4, 0.3, 0.1, center=(0.2, -0.1))
    fcs = ilm.ForcingModelAndRegion([ilm.AreaForcingModel(shp, ilm.RigidTransform((0.1, 0.0), 0.3), lambda s, T, t, fr, pp: s.fill(1.0)),
                                     ilm.LineForcingModel(shp, ilm.RigidTransform(), lambda s, T, t, fr, pp: s.fill(-2.0)),
                                     ilm.AreaForcingModel(lambda s, T, t, fr, pp: s.fill(0.5)),
                                     ilm.PointForcingModel((np.array([-1.9, 0.5]), np.array([0.5, 1.7])),
                                                           lambda s, T, t, fr, pp: s.fill(1.0), ddftype="m4prime")], scc)
    ilm.apply_forcing(scc.zeros_grid(), scc.zeros_grid(), None, 0.0, fcs, None, None, scc)
    fcv = ilm.ForcingModelAndRegion([ilm.AreaForcingModel(shp, ilm.RigidTransform(), lambda s, T, t, fr, pp: s.fill(1.0)),
                                     ilm.LineForcingModel(shp, ilm.RigidTransform(), lambda s, T, t, fr, pp: s.fill(-2.0))], vcc)
    ilm.apply_forcing(vcc.zeros_grid(), vcc.zeros_grid(), None, 0.0, fcv, None, None, vcc)
    ww, dd, dvv = vcc.zeros_gridcurl().fill(1.0), vcc.zeros_griddiv().fill(-1.0), vcc.zeros_surface().fill(0.3)
    vv = vcc.zeros_grid()
    ilm.vecfield_helmholtz(vv, ww, dd, dvv, (1.0, 0.5), vcc)
    ilm.masked_curlv_from_curlv_masked(ww, ww, dvv, vcc)
    ilm.divv_masked_from_masked_divv(vcc.zeros_griddiv(), dd, dvv, vcc)
    ilm.vectorpotential_from_curlv(vcc.zeros_gridcurl(), ww, vcc)
    ilm.scalarpotential_from_masked_divv(vcc.zeros_griddiv(), dd, dvv, vcc)
print("done (forcing, helmholtz)")

"""GPU parity of the forcing regions (src/forcing.jl) against the oracle.  Mirrors the reference's
"Forcing regions" testset (test/surface_ops.jl:445-542): an area heater, a line heater, a
temperature-dependent area heater on a rotated polygon and two M4' point sources on the cache of a body.
Line and point contributions are bit-exact (the gather adds in the reference's CSC order); the area
contribution is bit-exact given the mask, and the mask itself passes through the FFT (1e-12 norm-wise)."""
import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L
from ilm_b200 import forcing as F

pytestmark = pytest.mark.gpu

N = 128


def relerr(a, b):
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


def host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def triangle(ds):
    """Polygon([0,1,0.5],[0,0,0.5],ds) surrogate: midpoints of the subdivided sides, counter-clockwise."""
    vx, vy = np.array([0.0, 1.0, 0.5, 0.0]), np.array([0.0, 0.0, 0.5, 0.0])
    xs, ys = [], []
    for k in range(3):
        n = max(int(np.round(np.hypot(vx[k + 1] - vx[k], vy[k + 1] - vy[k]) / ds)), 1)
        t = np.linspace(0.0, 1.0, n + 1)[:-1]
        xs.append(vx[k] + t * (vx[k + 1] - vx[k]))
        ys.append(vy[k] + t * (vy[k + 1] - vy[k]))
    xv, yv = np.concatenate(xs + [[0.0]]), np.concatenate(ys + [[0.0]])
    return ilm.bodies._polygon_midpoints(xv, yv, (0.0, 0.0))


@pytest.fixture(scope="module", params=[False, True], ids=["host", "device"])
def setup(request):
    g = ilm.PhysicalGrid.centered(N)
    G = ilm.lgf.lgf_table(N)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    scache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=request.param)
    og = o.Grid(g.NX, g.NY, g.dx, g.I0)
    return g, G, og, scache


PHYS = {"areaheater1_flux": 3.0, "lineheater_flux": -2.0, "areaheater2_temp": 1.5, "areaheater2_coeff": 2.0}
PTS = (np.array([-1.2, 0.5]), np.array([0.5, 0.5]))
STR = [1.0, -1.0]


def _models(g):
    ds = 1.4 * g.dx

    def model1(sig, T, t, fr, pp):
        assert isinstance(fr, ilm.AreaRegionCache)
        sig.fill(pp["areaheater1_flux"])

    def model2(sig, T, t, fr, pp):
        assert isinstance(fr, ilm.LineRegionCache)
        sig.fill(pp["lineheater_flux"])

    def model3(sig, T, t, fr, pp):
        sig.set(pp["areaheater2_coeff"] * (pp["areaheater2_temp"] - T.numpy()))

    def model4(sig, T, t, fr, pp):
        assert isinstance(fr, ilm.PointRegionCache)
        sig.set(np.asarray(STR, dtype=float))

    afm = ilm.AreaForcingModel(ilm.bodies.circle(0.2, ds), ilm.RigidTransform((0.5, 0.0), 0.0), model1)
    lfm = ilm.LineForcingModel(ilm.bodies.rectangle(0.25, 0.25, ds), ilm.RigidTransform((0.0, 1.0), 0.0), model2)
    afm2 = ilm.AreaForcingModel(triangle(ds), ilm.RigidTransform((-1.0, -1.0), np.pi / 4), model3)
    pfm = ilm.PointForcingModel(PTS, model4, ddftype="m4prime")
    return afm, lfm, afm2, pfm, model1, model4


def test_forcing_regions_match_oracle(setup):
    g, G, og, scache = setup
    afm, lfm, afm2, pfm, _, _ = _models(g)
    fcache = ilm.ForcingModelAndRegion([afm, lfm, afm2, pfm], scache)
    assert isinstance(fcache[0].region_cache.mask, ilm.Nodes)
    assert fcache[3].region_cache.cache.ddftype == "m4prime"
    T = scache.zeros_grid()
    T.set(np.random.default_rng(3).standard_normal(len(T)))
    dT = scache.zeros_grid()
    dT.fill(7.0)                                    # apply_forcing! zeroes its output first
    ilm.apply_forcing(dT, T, None, 0.0, fcache, PHYS, None, scache)
    got = dT.array().copy()

    # oracle: the same four regions
    b1 = ilm.RigidTransform((0.5, 0.0), 0.0)(ilm.bodies.circle(0.2, 1.4 * g.dx))
    b2 = ilm.RigidTransform((0.0, 1.0), 0.0)(ilm.bodies.rectangle(0.25, 0.25, 1.4 * g.dx))
    b3 = ilm.RigidTransform((-1.0, -1.0), np.pi / 4)(triangle(1.4 * g.dx))
    oc1, oc2, oc3 = (o.ScalarCache(og, *b[:5], G) for b in (b1, b2, b3))
    Th = T.array()
    m1, m3 = oc1.mask(), oc3.mask()
    ref = np.zeros_like(Th)
    ref = o.forcing_area(ref, np.full_like(Th, PHYS["areaheater1_flux"]), m1)
    ref = o.forcing_line(ref, oc2.tabs[o.PRIMAL], np.full(oc2.N, PHYS["lineheater_flux"]))
    ref = o.forcing_area(ref, PHYS["areaheater2_coeff"] * (PHYS["areaheater2_temp"] - Th), m3)
    ref = o.forcing_line(ref, o.point_collection_table(og, *PTS, o.PRIMAL, "m4prime"), np.asarray(STR))
    assert relerr(got, ref) < 1e-12

    # masks through the FFT: 1e-12; given the library's masks the whole composition is bit-exact
    gm1, gm3 = fcache[0].region_cache.mask.array(), fcache[2].region_cache.mask.array()
    assert relerr(gm1, m1) < 1e-12 and relerr(gm3, m3) < 1e-12
    ref2 = np.zeros_like(Th)
    ref2 = o.forcing_area(ref2, np.full_like(Th, PHYS["areaheater1_flux"]), gm1)
    ref2 = o.forcing_line(ref2, oc2.tabs[o.PRIMAL], np.full(oc2.N, PHYS["lineheater_flux"]))
    ref2 = o.forcing_area(ref2, PHYS["areaheater2_coeff"] * (PHYS["areaheater2_temp"] - Th), gm3)
    ref2 = o.forcing_line(ref2, o.point_collection_table(og, *PTS, o.PRIMAL, "m4prime"), np.asarray(STR))
    assert np.array_equal(got, ref2)

    # physics: the line heater deposits flux * perimeter, the points their net strength (zero)
    only_line = scache.zeros_grid()
    ilm.apply_forcing(only_line, T, None, 0.0, [fcache[1]], PHYS, None, scache)
    assert abs(only_line.numpy().sum() * g.dx ** 2 - PHYS["lineheater_flux"] * b2[4].sum()) < 1e-10
    only_area = scache.zeros_grid()
    ilm.apply_forcing(only_area, T, None, 0.0, fcache[0], PHYS, None, scache)
    assert abs(only_area.numpy().sum() * g.dx ** 2 - 3.0 * np.pi * 0.2 ** 2) < 3e-2


def test_point_function_and_dimension_mismatch(setup):
    """test/surface_ops.jl:503-524: a point function gives the same forcing as fixed points; a strength vector of
    the wrong length throws DimensionMismatch."""
    global STR
    g, G, og, scache = setup
    _, _, _, pfm, _, model4 = _models(g)

    def point_function(T, t, fr, pp):
        return np.array([-1.2, 0.5]), np.array([0.5, 0.5])

    pfm2 = ilm.PointForcingModel(point_function, model4, ddftype="m4prime")
    fc, fc2 = ilm.ForcingModelAndRegion(pfm, scache), ilm.ForcingModelAndRegion(pfm2, scache)
    T, dT, dT2 = scache.zeros_grid(), scache.zeros_grid(), scache.zeros_grid()
    ilm.apply_forcing(dT, T, None, 0.0, fc, PHYS, None, scache)
    ilm.apply_forcing(dT2, T, None, 0.0, fc2, PHYS, None, scache)
    assert np.array_equal(dT.numpy(), dT2.numpy())
    ref = o.forcing_line(np.zeros(dT.shape), o.point_collection_table(og, *PTS, o.PRIMAL, "m4prime"), np.asarray(STR))
    assert np.array_equal(dT.array(), ref)
    assert abs(np.abs(dT.numpy()).max()) > 0
    old = STR
    try:
        STR = [1.0, -1.0, 2.0]
        with pytest.raises(ilm.DimensionMismatch):
            ilm.apply_forcing(dT2, T, None, 0.0, fc2, PHYS, None, scache)
    finally:
        STR = old


def test_whole_domain_area_forcing_and_generated_field(setup):
    """test/surface_ops.jl:526-540: area forcing without a shape; a spatial field used as the strength comes back
    unchanged (dT == generated_field())."""
    g, G, og, scache = setup
    _, _, _, _, model1, _ = _models(g)
    T, dT = scache.zeros_grid(), scache.zeros_grid()
    fc = ilm.ForcingModelAndRegion(ilm.AreaForcingModel(model1), scache)
    ilm.apply_forcing(dT, T, None, 0.0, fc, PHYS, None, scache)
    assert np.array_equal(dT.numpy(), np.full(len(dT), 3.0))
    assert np.array_equal(fc[0].region_cache.mask.numpy(), np.ones(len(dT)))

    def model6(sig, T, t, fr, pp):
        sig.set(fr.generated_field().numpy())

    afm = ilm.AreaForcingModel(model6, spatialfield=ilm.SpatialGaussian(0.5, 0.1, 0, 0, 10))
    fc = ilm.ForcingModelAndRegion(afm, scache)
    ilm.apply_forcing(dT, T, None, 0.0, fc, PHYS, None, scache)
    gf = fc[0].region_cache.generated_field().numpy()
    assert np.array_equal(dT.numpy(), gf)
    assert abs(gf.sum() * g.dx ** 2 - 10.0) < 1e-6              # the Gaussian integrates to its amplitude


def test_vector_cache_forcing(setup):
    """Area and line forcing with Edges grid data (SurfaceVectorCache regions, src/forcing.jl:197-232)."""
    g, G, og, scache = setup
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    vcache = ilm.SurfaceVectorCache(body, g, lgf_table=G, device=scache.device)
    shape = ilm.RigidTransform((0.3, -0.2), 0.4)(ilm.bodies.rectangle(0.4, 0.2, 1.4 * g.dx))
    rng = np.random.default_rng(5)
    su = rng.standard_normal(2 * shape[0].shape[0])

    def area_model(sig, v, t, fr, pp):
        sig.set(2.0 * v.numpy() + 1.0)

    def line_model(sig, v, t, fr, pp):
        sig.set(su)

    fc = ilm.ForcingModelAndRegion([ilm.AreaForcingModel(shape, ilm.RigidTransform(), area_model),
                                    ilm.LineForcingModel(shape, ilm.RigidTransform(), line_model)], vcache)
    v = vcache.zeros_grid()
    v.set(rng.standard_normal(len(v)))
    dv = vcache.zeros_grid()
    ilm.apply_forcing(dv, v, None, 0.0, fc, None, None, vcache)
    oc = o.VectorCache(og, *shape[:5], G)
    mu, mv = oc.mask_edges()
    n = shape[0].shape[0]
    ru = o.forcing_line(o.forcing_area(np.zeros_like(mu), 2.0 * v.u + 1.0, mu), oc.tabs[o.XEDGE], su[:n])
    rv = o.forcing_line(o.forcing_area(np.zeros_like(mv), 2.0 * v.v + 1.0, mv), oc.tabs[o.YEDGE], su[n:])
    assert relerr(dv.u, ru) < 1e-12 and relerr(dv.v, rv) < 1e-12
    gm = fc[0].region_cache.mask
    ru2 = o.forcing_line(o.forcing_area(np.zeros_like(mu), 2.0 * v.u + 1.0, gm.u), oc.tabs[o.XEDGE], su[:n])
    rv2 = o.forcing_line(o.forcing_area(np.zeros_like(mv), 2.0 * v.v + 1.0, gm.v), oc.tabs[o.YEDGE], su[n:])
    assert np.array_equal(dv.u, ru2) and np.array_equal(dv.v, rv2)


def test_shared_plan_is_a_full_cache(setup):
    """A child plan (ilm_plan_create_shared) gives the same Schur complement as a stand-alone cache on the same
    points, and leaves the parent's results unchanged."""
    g, G, og, scache = setup
    shape = ilm.bodies.ellipse(0.6, 0.3, 1.4 * g.dx, center=(0.2, 0.1))
    child = ilm.SurfaceScalarCache(shape, g, device=scache.device, parent=scache)
    alone = ilm.SurfaceScalarCache(shape, g, lgf_table=G, device=scache.device)
    S0 = host(ilm.create_RTLinvR(scache)).copy()
    assert np.array_equal(host(ilm.create_RTLinvR(child)), host(ilm.create_RTLinvR(alone)))
    assert np.array_equal(host(ilm.create_RTLinvR(scache)), S0)
    child.close()
    assert np.array_equal(host(ilm.create_RTLinvR(scache)), S0)


def test_forcing_abi_errors(setup):
    g, G, og, scache = setup
    lib = L.load()
    dT = scache.zeros_grid()
    with pytest.raises(ilm.MethodError):
        L.check(lib.ilm_forcing_area_add(scache._plan, 9, ilm.api._ptr(dT.data), None, ilm.api._ptr(dT.data)))
    with pytest.raises(ilm.MethodError):
        L.check(lib.ilm_forcing_line_add(scache._plan, -1, ilm.api._ptr(dT.data), ilm.api._ptr(dT.data)))
    with pytest.raises(ilm.MethodError):
        L.check(lib.ilm_forcing_area_add(scache._plan, L.NODES_PRIMAL, None, None, ilm.api._ptr(dT.data)))

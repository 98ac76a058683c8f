"""GPU parity of the vector-data cache (SurfaceVectorCache: VectorData / TensorData /
EdgeGradient operators and the 2N x 2N Schur builders) against the oracle."""
import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make_vcase(NX, NY, dx, I0, body):
    g = ilm.PhysicalGrid(NX, NY, dx, I0)
    G = ilm.lgf.lgf_table(max(NX, NY))
    cache = ilm.SurfaceVectorCache(body, g, lgf_table=G)
    oc = o.VectorCache(o.Grid(NX, NY, dx, I0), *body[:5], G)
    return cache, oc


@pytest.fixture(scope="module", params=["c1", "odd"])
def vc(request):
    if request.param == "c1":
        g = ilm.PhysicalGrid.centered(128)
        return make_vcase(g.NX, g.NY, g.dx, g.I0, ilm.bodies.circle(1.0, 1.4 * g.dx))
    return make_vcase(91, 67, 0.05, (44, 35), ilm.bodies.rectangle(0.6, 0.35, 0.07, center=(0.13, -0.21)))


def eg_arrays(q):
    return [q.component(i) for i in range(4)]


def test_tensor_regularize_interpolate(vc):
    cache, oc = vc
    N = cache.N
    rng = np.random.default_rng(31)
    v = rng.standard_normal(2 * N)
    vd = ilm.VectorData(N, data=v.copy())
    q = cache.zeros_gridgrad()
    ilm.regularize_normal(q, vd, cache)
    for a, b in zip(eg_arrays(q), oc.regularize_normal_v(v[:N], v[N:])):
        assert np.array_equal(a, b)
    ilm.regularize_normal_symm(q, vd, cache)
    for a, b in zip(eg_arrays(q), oc.regularize_normal_symm(v[:N], v[N:])):
        assert np.array_equal(a, b)
    A = [rng.standard_normal(s) for s in q.shapes]
    q.set_components(A)
    out = cache.zeros_surface()
    ilm.normal_interpolate(out, q, cache)
    tu, tv = oc.normal_interpolate_v(A)
    assert relerr(out.u, tu) < RTOL and relerr(out.v, tv) < RTOL
    ilm.normal_interpolate_symm(out, q, cache)
    tu, tv = oc.normal_interpolate_symm(A)
    assert relerr(out.u, tu) < RTOL and relerr(out.v, tv) < RTOL


def test_vector_scalar_bridges(vc):
    cache, oc = vc
    N, g = cache.N, cache.g
    rng = np.random.default_rng(32)
    v = rng.standard_normal(2 * N)
    vd = ilm.VectorData(N, data=v.copy())
    w = cache.zeros_gridcurl()
    ilm.regularize_normal_cross(w, vd, cache)
    assert np.array_equal(w.array(), oc.regularize_normal_cross_v(v[:N], v[N:]))
    f = cache.zeros_griddiv()
    ilm.regularize_normal_dot(f, vd, cache)
    assert np.array_equal(f.array(), oc.regularize_normal_dot_v(v[:N], v[N:]))
    tau = rng.standard_normal(4 * N)
    e = cache.zeros_grid()
    ilm.regularize_normal_dot(e, ilm.TensorData(N, data=tau.copy()), cache)
    ru, rv = oc.regularize_normal_dot_t([tau[i * N:(i + 1) * N] for i in range(4)])
    assert np.array_equal(e.u, ru) and np.array_equal(e.v, rv)
    psi = rng.standard_normal(g.layout_shape(L.NODES_DUAL))
    phi = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    out = cache.zeros_surface()
    ilm.normal_cross_interpolate(out, ilm.Nodes(ilm.Dual, g).set(psi), cache)
    a, b = oc.normal_cross_interpolate_v(psi)
    assert relerr(out.u, a) < RTOL and relerr(out.v, b) < RTOL
    ilm.normal_dot_interpolate(out, ilm.Nodes(ilm.Primal, g).set(phi), cache)
    a, b = oc.normal_dot_interpolate_v(phi)
    assert relerr(out.u, a) < RTOL and relerr(out.v, b) < RTOL
    u = rng.standard_normal(g.layout_shape(L.XEDGES)); vv = rng.standard_normal(g.layout_shape(L.YEDGES))
    e.set(np.concatenate([u.ravel(order="F"), vv.ravel(order="F")]))
    td = cache.zeros_surfacetensor()
    ilm.normal_dot_interpolate(td, e, cache)
    for i, ref in enumerate(oc.normal_dot_interpolate_t(u, vv)):
        assert relerr(td.component(i), ref) < RTOL


def test_tensor_stencils_bit_exact(vc):
    cache, oc = vc
    g = cache.g
    rng = np.random.default_rng(33)
    u = rng.standard_normal(g.layout_shape(L.XEDGES)); v = rng.standard_normal(g.layout_shape(L.YEDGES))
    e = cache.zeros_grid().set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    q = cache.zeros_gridgrad().fill(5.0)
    ilm.grad(q, e, cache)
    for a, b in zip(eg_arrays(q), o.grad_e2t(oc.grid, u, v)):
        assert np.array_equal(a, b / g.dx)
    A = [rng.standard_normal(s) for s in q.shapes]
    q.set_components(A)
    e.fill(5.0)
    ilm.divergence(e, q, cache)
    du, dv = o.divergence_t2e(oc.grid, *A)
    assert np.array_equal(e.u, du / g.dx) and np.array_equal(e.v, dv / g.dx)


def test_vector_composites(vc):
    cache, oc = vc
    N, g = cache.N, cache.g
    rng = np.random.default_rng(34)
    v = rng.standard_normal(2 * N)
    vd = ilm.VectorData(N, data=v.copy())
    e = cache.zeros_grid()
    for symm, fn in ((False, ilm.surface_divergence), (True, ilm.surface_divergence_symm)):
        fn(e, vd, cache)
        du, dv = oc.surface_divergence_v(v[:N], v[N:], symm)
        assert np.array_equal(e.u, du) and np.array_equal(e.v, dv)
    u = rng.standard_normal(g.layout_shape(L.XEDGES)); vv = rng.standard_normal(g.layout_shape(L.YEDGES))
    e.set(np.concatenate([u.ravel(order="F"), vv.ravel(order="F")]))
    out = cache.zeros_surface()
    for symm, fn in ((False, ilm.surface_grad), (True, ilm.surface_grad_symm)):
        fn(out, e, cache)
        tu, tv = oc.surface_grad_v(u, vv, symm)
        assert relerr(out.u, tu) < RTOL and relerr(out.v, tv) < RTOL
    w = cache.zeros_gridcurl()
    ilm.surface_curl(w, vd, cache)
    assert np.array_equal(w.array(), oc.surface_curl_v2n(v[:N], v[N:]))
    psi = rng.standard_normal(g.layout_shape(L.NODES_DUAL))
    ilm.surface_curl(out, ilm.Nodes(ilm.Dual, g).set(psi), cache)
    a, b = oc.surface_curl_n2v(psi)
    assert relerr(out.u, a) < RTOL and relerr(out.v, b) < RTOL
    # scalar-data forms dispatch on a vector cache exactly as on a scalar cache (src/surface_operators.jl:363-366)
    f = rng.standard_normal(N)
    th = cache.zeros_griddiv()
    ilm.surface_divergence(th, ilm.ScalarData(N, data=f.copy()), cache)
    assert np.array_equal(th.array(), oc.surface_divergence(f))


def test_vector_reference_sequence():
    """test/surface_ops.jl:124-137: symm round trip extrema ~ +-22.5 at dx = 0.04."""
    dx, NX = 0.04, 104
    body = ilm.bodies.circle(1.0, 1.4 * dx)
    cache, _ = make_vcase(NX, NX, dx, (NX // 2, NX // 2), body)
    N = cache.N
    th = 2 * np.pi * np.arange(N) / N
    vs = cache.zeros_surface().set(np.concatenate([np.sin(th - np.pi / 4), np.zeros(N)]))
    dq = cache.zeros_gridgrad()
    ilm.regularize_normal(dq, vs, cache)
    ilm.normal_interpolate(vs, dq, cache)
    dq.fill(0.0)
    vs.u[:] = np.sin(th - np.pi / 4)
    ilm.regularize_normal_symm(dq, vs, cache)
    ilm.normal_interpolate_symm(vs, dq, cache)
    assert abs(vs.u.max() - 22.5) < 1.0 and abs(vs.u.min() + 22.5) < 1.0
    A = ilm.create_GLinvD(cache, scale=dx)
    assert A.shape == (2 * N, 2 * N)
    assert abs(np.abs(np.linalg.eigvals(A)).max() - 0.45) < 0.1                    # :145-146


def test_vector_schur_matrices():
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    cache, oc = make_vcase(g.NX, g.NY, g.dx, g.I0, body)
    dx, N = g.dx, cache.N
    assert relerr(ilm.create_RTLinvR(cache), oc.create_RTLinvR_v()) < RTOL
    assert relerr(ilm.create_CLinvCT(cache, scale=dx), oc.create_CLinvCT_v(scale=dx)) < RTOL
    assert relerr(ilm.create_CL2invCT(cache, scale=dx, cols=(3, 20)), oc.create_CL2invCT(scale=dx, cols=range(3, 20))) < RTOL
    assert relerr(ilm.create_GLinvD(cache, scale=dx), oc.create_GLinvD_v(scale=dx)) < RTOL
    assert relerr(ilm.create_GLinvD_symm(cache, scale=dx, cols=(N - 5, N + 9)),
                  oc.create_GLinvD_v(scale=dx, symm=True, cols=range(N - 5, N + 9))) < RTOL
    assert relerr(ilm.create_CLinvCT_scalar(cache, scale=dx), oc.create_CLinvCT_scalar(scale=dx)) < RTOL
    assert relerr(ilm.create_nRTRn(cache), oc.create_nRTRn_v()) < RTOL
    m = ilm.mask(cache)
    mu, mv = oc.mask_edges()
    assert relerr(m.u, mu) < RTOL and relerr(m.v, mv) < RTOL
    # scalar caches reject the vector-only builders (MethodError in the reference)
    sc = ilm.SurfaceScalarCache(body, g)
    with pytest.raises(ilm.MethodError):
        ilm.create_CL2invCT(sc)
    with pytest.raises(ilm.MethodError):
        ilm.surface_divergence(sc.zeros_gridgrad(), ilm.VectorData(sc.N), sc)


@pytest.mark.parametrize("complementary", [False, True])
def test_vector_mask_products(vc, complementary):
    """mask!/complementary_mask! on a vector cache: Edges, Nodes{Dual}, Nodes{Primal}, EdgeGradient
    (src/surface_operators.jl:796-800, 819-823, 904-923; test/surface_ops.jl:212-232)."""
    cache, oc = vc
    rng = np.random.default_rng(17)
    fn = ilm.complementary_mask if complementary else ilm.mask
    q = cache.zeros_grid()
    au, av = rng.standard_normal(q.ushape), rng.standard_normal(q.vshape)
    q.set(np.concatenate([au.ravel(order="F"), av.ravel(order="F")]))
    fn(q, cache)
    ru, rv = oc.mask_product_v((au, av), "edges", complementary)
    assert relerr(q.u, ru) < RTOL and relerr(q.v, rv) < RTOL
    for celltype, kind in ((ilm.Primal, o.PRIMAL), (ilm.Dual, o.DUAL)):
        w = ilm.Nodes(celltype, cache.g)
        a = rng.standard_normal(w.shape)
        w.set(a)
        fn(w, cache)
        assert relerr(w.array(), oc.mask_product_v(a, kind, complementary)) < RTOL
    t = cache.zeros_gridgrad()
    A = [rng.standard_normal(s) for s in t.shapes]
    t.set_components(A)
    fn(t, cache)
    for got, ref in zip(eg_arrays(t), oc.mask_product_v(tuple(A), "edgegrad", complementary)):
        assert relerr(got, ref) < RTOL
    with pytest.raises(ilm.MethodError):
        fn(ilm.XEdges(cache.g), cache)


def test_stokes_flow_matches_oracle():
    """test/literate/stokes.jl:98-166 (config C5a scaled down): rectangle translating at unit speed."""
    g = ilm.PhysicalGrid.centered(128)
    body = ilm.bodies.rectangle(0.5, 0.25, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(128)
    cache = ilm.SurfaceVectorCache(body, g, lgf_table=G)
    oc = o.VectorCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G)
    N = cache.N
    vplus = np.concatenate([np.ones(N), np.zeros(N)])          # stokes.jl:186-191
    v, s, sigma, S, Ss = ilm.stokes_flow(cache, vplus)
    vu, vv, sr, sigr = o.stokes_solve(oc, vplus)
    assert relerr(v.u, vu) < 1e-9 and relerr(v.v, vv) < 1e-9
    assert relerr(s.array(), sr) < 1e-9
    assert relerr(sigma.data, sigr) < 1e-5                      # solution of an ill-conditioned system
    # the factorisations are reusable for other boundary data
    vplus2 = np.concatenate([np.zeros(N), np.ones(N)])
    v2 = ilm.stokes_flow(cache, vplus2, S=S, Ss=Ss)[0]
    vu2, vv2, _, _ = o.stokes_solve(oc, vplus2)
    assert relerr(v2.u, vu2) < 1e-9 and relerr(v2.v, vv2) < 1e-9
    # the constraint: the velocity interpolated on the surface is the prescribed average (v+ + v-)/2
    vb = cache.zeros_surface()
    ilm.interpolate(vb, v, cache)
    assert np.abs(vb.u - 0.5).max() < 1e-8 and np.abs(vb.v).max() < 1e-8

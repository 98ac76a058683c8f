"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
identical seeded inputs.  Bit-exact for index/weight tables, regularization
and the difference stencils; 1e-12 (norm-wise, relative to the field's max
magnitude -- SURVEY.md section 7 "hard parts") for every floating-point field
that goes through a reduction tree or the FFT."""
import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L

pytestmark = pytest.mark.gpu

RTOL = 1e-12
KIND = {L.NODES_PRIMAL: o.PRIMAL, L.NODES_DUAL: o.DUAL, L.XEDGES: o.XEDGE, L.YEDGES: o.YEDGE}


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a - b).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 1.0
    return d / (s if s > 0 else 1.0)


def make_case(NX, NY, dx, I0, body, ddf="yang3", scaling=ilm.GridScaling, device=False, lgf_rule="accurate"):
    g = ilm.PhysicalGrid(NX, NY, dx, I0)
    G = ilm.lgf.lgf_table(max(NX, NY), rule=lgf_rule)
    cache = ilm.SurfaceScalarCache(body, g, scaling=scaling, ddftype=ddf, lgf_table=G, device=device)
    og = o.Grid(NX, NY, dx, I0)
    oc = o.ScalarCache(og, *body[:5], G, ddf=ddf, scaling=o.GRID_SCALING if scaling == ilm.GridScaling else o.INDEX_SCALING)
    return cache, oc


@pytest.fixture(scope="module")
def c1():
    """BASELINE config C1: Dirichlet Poisson on a circle, 128x128, Yang3."""
    g = ilm.PhysicalGrid.centered(128)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    return make_case(g.NX, g.NY, g.dx, g.I0, body)


@pytest.fixture(scope="module")
def odd():
    """Non-square, odd-sized grid with an off-centre ellipse (ragged rows, misaligned layouts)."""
    body = ilm.bodies.ellipse(0.9, 0.45, 0.07, center=(0.13, -0.21))
    return make_case(91, 67, 0.05, (44, 35), body)


# ---------------------------------------------------------------- tables
@pytest.mark.parametrize("ddf", ["yang3", "m3", "roma", "m4prime", "witchhat"])
@pytest.mark.parametrize("scaling", [ilm.GridScaling, ilm.IndexScaling])
def test_tables_bit_exact(ddf, scaling):
    body = ilm.bodies.ellipse(0.9, 0.45, 0.07, center=(0.13, -0.21))
    cache, oc = make_case(91, 67, 0.05, (44, 35), body, ddf=ddf, scaling=scaling)
    for layout, kind in KIND.items():
        idx, wR, wE = cache.table(layout)
        tab = oc.tabs[kind]
        lin = np.transpose(tab.linear_index(), (0, 2, 1))        # oracle [k][a][b] -> [k][b][a]
        assert np.array_equal(idx, lin)
        assert np.array_equal(wR, np.transpose(tab.wR, (0, 2, 1)))
        assert np.array_equal(wE, np.transpose(tab.wE, (0, 2, 1)))


def test_tables_clipped_at_boundary():
    """Points next to the grid edge: window entries outside the field are dropped."""
    g = ilm.PhysicalGrid(40, 36, 0.1, (3, 4))
    body = ilm.bodies.circle(0.25, 0.14)
    cache, oc = make_case(g.NX, g.NY, g.dx, g.I0, body)
    for layout, kind in KIND.items():
        idx, wR, wE = cache.table(layout)
        lin = np.transpose(oc.tabs[kind].linear_index(), (0, 2, 1))
        assert (lin < 0).any()
        assert np.array_equal(idx, lin)
        assert np.array_equal(wR, np.transpose(oc.tabs[kind].wR, (0, 2, 1)))
    f = np.random.default_rng(0).standard_normal(cache.N)
    s = cache.zeros_grid()
    ilm.regularize(s, ilm.ScalarData(cache.N, data=f), cache)
    assert np.array_equal(s.array(), oc.regularize(f))


# ---------------------------------------------------------------- regularize / interpolate
@pytest.mark.parametrize("case", ["c1", "odd"])
def test_regularize_interpolate(case, request):
    cache, oc = request.getfixturevalue(case)
    rng = np.random.default_rng(1)
    f = rng.standard_normal(cache.N)
    fd = ilm.ScalarData(cache.N, data=f.copy())
    for celltype, kind in ((ilm.Primal, o.PRIMAL), (ilm.Dual, o.DUAL)):
        s = ilm.Nodes(celltype, cache.g)
        ilm.regularize(s, fd, cache)
        ref = o.regularize(oc.tabs[kind], f)
        assert np.array_equal(s.array(), ref)                       # deterministic gather: bit-exact
        w = rng.standard_normal(ref.shape)
        sw = ilm.Nodes(celltype, cache.g).set(w)
        out = cache.zeros_surface()
        ilm.interpolate(out, sw, cache)
        assert relerr(out.data, o.interpolate(oc.tabs[kind], w)) < RTOL
    # Edges <-> VectorData
    fv = rng.standard_normal(2 * cache.N)
    q = cache.zeros_gridgrad()
    ilm.regularize(q, ilm.VectorData(cache.N, data=fv.copy()), cache)
    ru, rv = oc.regularize_edges(fv[:cache.N], fv[cache.N:])
    assert np.array_equal(q.u, ru) and np.array_equal(q.v, rv)
    q.set(np.concatenate([rng.standard_normal(ru.shape).ravel(order="F"), rng.standard_normal(rv.shape).ravel(order="F")]))
    vb = cache.zeros_surfacevec()
    ilm.interpolate(vb, q, cache)
    su, sv = oc.interpolate_edges(q.u, q.v)
    assert relerr(vb.u, su) < RTOL and relerr(vb.v, sv) < RTOL


@pytest.mark.parametrize("case", ["c1", "odd"])
def test_normal_variants(case, request):
    cache, oc = request.getfixturevalue(case)
    rng = np.random.default_rng(2)
    f = rng.standard_normal(cache.N)
    fd = ilm.ScalarData(cache.N, data=f.copy())
    q = cache.zeros_gridgrad()
    ilm.regularize_normal(q, fd, cache)
    ru, rv = oc.regularize_normal(f)
    assert np.array_equal(q.u, ru) and np.array_equal(q.v, rv)
    ilm.regularize_normal_cross(q, fd, cache)
    cu, cv = oc.regularize_normal_cross(f)
    assert np.array_equal(q.u, cu) and np.array_equal(q.v, cv)
    u = rng.standard_normal(ru.shape)
    v = rng.standard_normal(rv.shape)
    q.set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    vn = cache.zeros_surface()
    ilm.normal_interpolate(vn, q, cache)
    assert relerr(vn.data, oc.normal_interpolate(u, v)) < RTOL
    ilm.normal_cross_interpolate(vn, q, cache)
    assert relerr(vn.data, oc.normal_cross_interpolate(u, v)) < RTOL


# ---------------------------------------------------------------- stencils
@pytest.mark.parametrize("case", ["c1", "odd"])
def test_stencils_bit_exact(case, request):
    cache, oc = request.getfixturevalue(case)
    g, og = cache.g, oc.grid
    rng = np.random.default_rng(3)
    u = rng.standard_normal(g.layout_shape(L.XEDGES))
    v = rng.standard_normal(g.layout_shape(L.YEDGES))
    p = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    s = rng.standard_normal(g.layout_shape(L.NODES_DUAL))
    q = ilm.Edges(g).set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    pn = ilm.Nodes(ilm.Primal, g).set(p)
    sn = ilm.Nodes(ilm.Dual, g).set(s)
    out_p = ilm.Nodes(ilm.Primal, g).fill(7.0)        # outputs are fully overwritten
    ilm.divergence(out_p, q, cache)
    assert np.array_equal(out_p.array(), oc.divergence(u, v))
    out_q = ilm.Edges(g).fill(7.0)
    ilm.grad(out_q, pn, cache)
    gu, gv = oc.grad(p)
    assert np.array_equal(out_q.u, gu) and np.array_equal(out_q.v, gv)
    ilm.curl(out_q, sn, cache)
    cu, cv = oc.curl_n2e(s)
    assert np.array_equal(out_q.u, cu) and np.array_equal(out_q.v, cv)
    out_s = ilm.Nodes(ilm.Dual, g).fill(7.0)
    ilm.curl(out_s, q, cache)
    assert np.array_equal(out_s.array(), oc.curl_e2n(u, v))
    ilm.laplacian(out_s, sn, cache)
    assert np.array_equal(out_s.array(), oc.laplacian(s, o.DUAL))
    ilm.laplacian(out_p, pn, cache)
    assert np.array_equal(out_p.array(), oc.laplacian(p, o.PRIMAL))
    ilm.laplacian(out_q, q, cache)
    assert np.array_equal(out_q.u, oc.laplacian(u, o.XEDGE)) and np.array_equal(out_q.v, oc.laplacian(v, o.YEDGE))


# ---------------------------------------------------------------- inverse Laplacian
@pytest.mark.parametrize("case", ["c1", "odd"])
def test_inverse_laplacian(case, request):
    cache, oc = request.getfixturevalue(case)
    g = cache.g
    rng = np.random.default_rng(4)
    for celltype, kind in ((ilm.Primal, o.PRIMAL), (ilm.Dual, o.DUAL)):
        w = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL if celltype == ilm.Primal else L.NODES_DUAL))
        wn = ilm.Nodes(celltype, g).set(w)
        ilm.inverse_laplacian(wn, cache)
        assert relerr(wn.array(), oc.inverse_laplacian(w)) < RTOL
    u = rng.standard_normal(g.layout_shape(L.XEDGES))
    v = rng.standard_normal(g.layout_shape(L.YEDGES))
    q = ilm.Edges(g).set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    ilm.inverse_laplacian(q, cache)        # u and v ride one complex transform
    assert relerr(q.u, oc.inverse_laplacian(u)) < RTOL and relerr(q.v, oc.inverse_laplacian(v)) < RTOL
    z = ilm.Nodes(ilm.Primal, g)
    ilm.inverse_laplacian(z, cache)        # iszero short-circuit of the reference: zeros stay zeros
    assert not np.any(z.array())


def test_inverse_laplacian_unit_impulses(c1):
    """A unit impulse returns the LGF table itself: (G - c0)/factor."""
    cache, oc = c1
    g = cache.g
    w = np.zeros(g.layout_shape(L.NODES_DUAL))
    w[5, 9] = 1.0
    wn = ilm.Nodes(ilm.Dual, g).set(w)
    ilm.inverse_laplacian(wn, cache)
    ii = np.abs(np.arange(g.NX) - 5)
    jj = np.abs(np.arange(g.NY) - 9)
    ref = (oc.lgf[np.ix_(ii, jj)] - oc.c0) / oc.factor
    assert relerr(wn.array(), ref) < RTOL


def test_intfact_kernel(c1):
    """plan_intfact on the same engine: E_a convolution, a = 0.5."""
    cache, oc = c1
    g = cache.g
    E = ilm.lgf.intfact_table(0.5, g.NX)
    kid = cache.add_kernel(E)
    assert kid == 1
    w = np.random.default_rng(5).standard_normal(g.layout_shape(L.NODES_DUAL))
    wn = ilm.Nodes(ilm.Dual, g).set(w)
    L.check(cache._lib.ilm_convolve(cache._plan, kid, L.NODES_DUAL, wn.data.ctypes.data))
    ref = o.ConvPlan(E[:g.NX, :g.NY]).apply(w)
    assert relerr(wn.array(), ref) < RTOL
    # a = 0 is the identity
    kid0 = cache.add_kernel(ilm.lgf.intfact_table(0.0, g.NX))
    wn.set(w)
    L.check(cache._lib.ilm_convolve(cache._plan, kid0, L.NODES_DUAL, wn.data.ctypes.data))
    assert relerr(wn.array(), w) < RTOL


# ---------------------------------------------------------------- composites, mask
@pytest.mark.parametrize("case", ["c1", "odd"])
def test_surface_grid_composites(case, request):
    cache, oc = request.getfixturevalue(case)
    g = cache.g
    rng = np.random.default_rng(6)
    f = rng.standard_normal(cache.N)
    fd = ilm.ScalarData(cache.N, data=f.copy())
    th = cache.zeros_grid()
    ilm.surface_divergence(th, fd, cache)
    assert np.array_equal(th.array(), oc.surface_divergence(f))
    ilm.surface_divergence_cross(th, fd, cache)
    assert np.array_equal(th.array(), oc.surface_divergence_cross(f))
    w = cache.zeros_gridcurl()
    ilm.surface_curl(w, fd, cache)
    assert np.array_equal(w.array(), oc.surface_curl_s2n(f))
    ilm.surface_curl_cross(w, fd, cache)
    assert np.array_equal(w.array(), oc.surface_curl_cross_s2n(f))
    phi = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    psi = rng.standard_normal(g.layout_shape(L.NODES_DUAL))
    vn = cache.zeros_surface()
    ilm.surface_grad(vn, ilm.Nodes(ilm.Primal, g).set(phi), cache)
    assert relerr(vn.data, oc.surface_grad(phi)) < RTOL
    ilm.surface_grad_cross(vn, ilm.Nodes(ilm.Primal, g).set(phi), cache)
    assert relerr(vn.data, oc.surface_grad_cross(phi)) < RTOL
    ilm.surface_curl(vn, ilm.Nodes(ilm.Dual, g).set(psi), cache)
    assert relerr(vn.data, oc.surface_curl_n2s(psi)) < RTOL
    ilm.surface_curl_cross(vn, ilm.Nodes(ilm.Dual, g).set(psi), cache)
    assert relerr(vn.data, oc.surface_curl_cross_n2s(psi)) < RTOL


def test_mask(c1):
    cache, oc = c1
    m = ilm.mask(cache)
    assert relerr(m.array(), oc.mask()) < RTOL
    cm = ilm.complementary_mask(cache)
    assert relerr(cm.array(), 1.0 - oc.mask()) < RTOL


@pytest.mark.parametrize("case", ["c1", "odd"])
@pytest.mark.parametrize("complementary", [False, True])
def test_mask_products(case, complementary, request):
    """mask!(w, cache) / complementary_mask!(w, cache) on every layout of a scalar cache
    (src/surface_operators.jl:788-823, 880-902; test/surface_ops.jl:190-210)."""
    cache, oc = request.getfixturevalue(case)
    rng = np.random.default_rng(5)
    fn = ilm.complementary_mask if complementary else ilm.mask
    mk = {o.PRIMAL: lambda: ilm.Nodes(ilm.Primal, cache.g), o.DUAL: lambda: ilm.Nodes(ilm.Dual, cache.g),
          o.XEDGE: lambda: ilm.XEdges(cache.g), o.YEDGE: lambda: ilm.YEdges(cache.g)}
    for kind, make in mk.items():
        w = make()
        a = rng.standard_normal(w.shape)
        w.set(a)
        assert fn(w, cache) is w
        assert relerr(w.array(), oc.mask_product(a, kind, complementary)) < RTOL
    q = cache.zeros_gridgrad()
    au, av = rng.standard_normal(q.ushape), rng.standard_normal(q.vshape)
    q.set(np.concatenate([au.ravel(order="F"), av.ravel(order="F")]))
    fn(q, cache)
    assert relerr(q.u, oc.mask_product(au, o.XEDGE, complementary)) < RTOL
    assert relerr(q.v, oc.mask_product(av, o.YEDGE, complementary)) < RTOL
    # ones -> area of the body (test/surface_ops.jl:193-207, 1e-3 on the reference's finer grid)
    if case == "c1" and not complementary:
        one = ilm.Nodes(ilm.Dual, cache.g).fill(1.0)
        ilm.mask(one, cache)
        area = o.dot_grid(oc.grid, one.array(), np.ones(one.shape), o.DUAL)
        assert abs(area - np.pi) < 2e-2
    with pytest.raises(ilm.MethodError):
        ilm.mask(ilm.EdgeGradient(cache.g), cache)


# ---------------------------------------------------------------- Schur builders, dense solve
def test_schur_matrices(c1):
    cache, oc = c1
    S = ilm.create_RTLinvR(cache)
    Sref = oc.create_RTLinvR()
    assert S.shape == Sref.shape
    assert relerr(S, Sref) < RTOL
    assert relerr(ilm.create_RTLinvR(cache, scale=2.5, cols=(3, 10)), 2.5 * Sref[:, 3:10]) < RTOL   # odd-sized block
    dx = cache.g.dx
    assert relerr(ilm.create_CLinvCT(cache, scale=dx), oc.create_CLinvCT(scale=dx)) < RTOL
    assert relerr(ilm.create_GLinvD(cache, scale=dx), oc.create_GLinvD(scale=dx)) < RTOL
    assert relerr(ilm.create_GLinvD_cross(cache, scale=dx, cols=(0, 16)), oc.create_GLinvD_cross(scale=dx, cols=range(16))) < RTOL
    assert relerr(ilm.create_nRTRn(cache), oc.create_nRTRn()) < RTOL
    assert relerr(ilm.create_surface_filter(cache), oc.create_surface_filter()) < RTOL


def test_schur_direct_table_form(c1):
    """the transform-free table form equals the column-solve path."""
    cache, oc = c1
    Sd = ilm.create_RTLinvR_direct(cache, scale=1.5)
    assert relerr(Sd, 1.5 * oc.create_RTLinvR()) < RTOL
    assert relerr(ilm.create_RTLinvR_direct(cache, cols=(5, 12)), ilm.create_RTLinvR(cache, cols=(5, 12))) < RTOL


def test_schur_reference_physics(c1):
    """The reference's own checks (test/surface_ops.jl:69-77) on the GPU matrices."""
    cache, _ = c1
    dx = cache.g.dx
    assert abs(np.abs(np.linalg.eigvals(ilm.create_CLinvCT(cache, scale=dx))).max() - 0.2) < 0.1
    assert abs(np.abs(np.linalg.eigvals(ilm.create_GLinvD(cache, scale=dx))).max() - 0.45) < 0.1
    # sigma_max(nRTRn) ~ 11 on the reference fixture (dx = 0.04) and scales like 1/dx
    assert abs(np.linalg.svd(ilm.create_nRTRn(cache), compute_uv=False).max() * dx / 0.04 - 11) < 1.5


@pytest.mark.parametrize("n", [1, 7, 32, 33, 141, 500, 2500, 4200])     # 4200: 16-CTA register panels, rank-128 updates
def test_dense_lu_solve(n):
    import scipy.linalg
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 0.1 * n * np.eye(n) * (rng.random(n) - 0.5)[:, None]
    b = rng.standard_normal(n)
    lu = ilm.LU(A)
    x = lu.solve(b)
    lu_ref, piv_ref = scipy.linalg.lu_factor(A)
    assert np.array_equal(lu.ipiv - 1, piv_ref)                     # same pivot sequence as LAPACK
    assert relerr(lu.lu.reshape((n, n), order="F"), lu_ref) < 1e-11
    xref = scipy.linalg.lu_solve((lu_ref, piv_ref), b)
    assert relerr(x, xref) < 1e-10 * max(1.0, np.linalg.cond(A) * 1e-3)
    assert np.abs(A @ x - b).max() < 1e-10 * np.abs(b).max() * n


@pytest.mark.parametrize("n", [64, 256, 1024])
def test_dense_lu_ties_go_to_the_first_row(n):
    """A Hadamard matrix: every pivot search is a tie of |a| over all remaining rows, and the elimination is exact in
    fp64 (dyadic entries), so LAPACK's choice -- the first maximum in the interchanged row order -- must be reproduced
    exactly although the panel kernel never moves rows (it tracks logical row indices), across warps and CTAs."""
    import scipy.linalg
    A = scipy.linalg.hadamard(n).astype(np.float64)
    lu = ilm.LU(A)
    lu_ref, piv_ref = scipy.linalg.lu_factor(A)
    assert np.array_equal(lu.ipiv - 1, piv_ref)
    assert np.array_equal(lu.lu.reshape((n, n), order="F"), lu_ref)
    b = np.arange(1.0, n + 1.0)
    assert relerr(lu.solve(b), scipy.linalg.lu_solve((lu_ref, piv_ref), b)) < 1e-13


def test_matvec_pow():
    rng = np.random.default_rng(8)
    n = 141
    Cm = rng.standard_normal((n, n)) / n
    s = rng.standard_normal(n)
    out = s.copy()
    ilm.matvec_pow(Cm, 5, out)
    assert relerr(out, np.linalg.matrix_power(Cm, 5) @ s) < RTOL


# ---------------------------------------------------------------- the north-star algorithm
def test_dirichlet_poisson_c1(c1):
    """test/literate/dirichlet.jl on config C1, GPU vs oracle, every field."""
    cache, oc = c1
    fplus = cache.points()[0].copy()
    f, s, S = ilm.dirichlet_poisson(cache, fplus)
    fr, sr, Sr = o.dirichlet_solve(oc, fplus)
    assert relerr(S, Sr) < RTOL
    # the multiplier s solves an ill-conditioned system: compare the residuals of
    # both solutions in the oracle's system and the solutions up to cond(S)*eps
    cond = np.linalg.cond(Sr)
    assert relerr(s.data, sr) < 50 * cond * np.finfo(float).eps
    assert relerr(f.array(), fr) < 1e-10
    # same S (the oracle's) -> the rest of the chain matches tightly
    f2, s2, _ = ilm.dirichlet_poisson(cache, fplus, S=Sr)
    assert relerr(s2.data, sr) < 50 * cond * np.finfo(float).eps
    # physics: zero inside, x/r^2 outside
    fa = f.array()
    i0 = cache.g.I0[0] - 1
    assert abs(fa[i0, i0]) < 1e-2


def test_device_resident_mode():
    """Torch CUDA tensors in, CUDA tensors out: no host staging."""
    import torch
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    cache, oc = make_case(g.NX, g.NY, g.dx, g.I0, body, device=True)
    f = np.random.default_rng(9).standard_normal(cache.N)
    fd = cache.zeros_surface().set(f)
    assert fd.data.is_cuda
    s = cache.zeros_grid()
    ilm.regularize(s, fd, cache)
    ilm.inverse_laplacian(s, cache)
    out = cache.zeros_surface()
    ilm.interpolate(out, s, cache)
    cache.sync()
    ref = oc.interpolate(oc.inverse_laplacian(oc.regularize(f)))
    assert relerr(out.numpy(), ref) < RTOL
    S = ilm.create_RTLinvR(cache)
    assert S.is_cuda
    assert relerr(S.cpu().numpy(), oc.create_RTLinvR()) < RTOL
    fsol, ssol, _ = ilm.dirichlet_poisson(cache, cache.points()[0].copy(), S=S)
    fr, sr, _ = o.dirichlet_solve(oc, oc.x.copy())
    assert relerr(fsol.numpy().reshape(fr.shape, order="F"), fr) < 1e-10
    assert cache.launch_count() > 0


def test_update_points_moving_body():
    """update_system (src/system.jl:26-50): new points, same Ghat."""
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(0.8, 1.4 * g.dx, center=(-0.3, 0.0))
    cache, _ = make_case(g.NX, g.NY, g.dx, g.I0, body)
    body2 = ilm.bodies.circle(0.8, 1.4 * g.dx, center=(0.25, 0.1))
    cache.update_points(body2)
    _, oc2 = make_case(g.NX, g.NY, g.dx, g.I0, body2)
    idx, wR, _ = cache.table(L.NODES_PRIMAL)
    assert np.array_equal(wR, np.transpose(oc2.tabs[o.PRIMAL].wR, (0, 2, 1)))
    assert relerr(ilm.create_RTLinvR(cache, cols=(0, 8)), oc2.create_RTLinvR(cols=range(8))) < RTOL


def test_zero_body_cache():
    """test/surface_ops.jl:356-371."""
    g = ilm.PhysicalGrid.centered(32)
    z = np.zeros(0)
    cache = ilm.SurfaceScalarCache((z, z, z, z, z), g)
    s = cache.zeros_grid().fill(3.0)
    ilm.regularize(s, cache.zeros_surface(), cache)
    assert not np.any(s.array())
    assert np.all(ilm.mask(cache).array() == 1.0)
    assert ilm.create_RTLinvR(cache).shape == (0, 0)


def test_error_behaviour(c1):
    """MethodError on container-type mismatch, DimensionMismatch on sizes (test/tools.jl:34-35,86)."""
    cache, _ = c1
    with pytest.raises(ilm.MethodError):
        ilm.regularize(cache.zeros_grid(), cache.zeros_surfacevec(), cache)
    with pytest.raises(ilm.MethodError):
        ilm.divergence(cache.zeros_gridcurl(), cache.zeros_gridgrad(), cache)
    with pytest.raises(ilm.DimensionMismatch):
        ilm.regularize(cache.zeros_grid(), ilm.ScalarData(cache.N + 1), cache)
    with pytest.raises(ilm.DimensionMismatch):
        ilm.create_RTLinvR(cache, cols=(0, cache.N + 1))
    with pytest.raises(ilm.DimensionMismatch):
        small = ilm.PhysicalGrid.centered(64)
        ilm.interpolate(cache.zeros_surface(), ilm.Nodes(ilm.Primal, small), cache)
    with pytest.raises(ilm.DimensionMismatch):
        ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 0.1), ilm.PhysicalGrid.centered(64),
                               lgf_table=ilm.lgf.lgf_table(16))


def test_reference_lgf_rule(c1):
    """The reference-compatible 100-node table runs through the same path."""
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    cache, oc = make_case(g.NX, g.NY, g.dx, g.I0, body, lgf_rule="gl100")
    assert relerr(ilm.create_RTLinvR(cache), oc.create_RTLinvR()) < RTOL


def test_neumann_poisson_c1(c1):
    """test/literate/neumann.jl:101-142 (BASELINE config C2 algorithm) GPU vs oracle, every output."""
    cache, oc = c1
    vnp = cache.normals()[0].copy()
    fr, dfr, sr, dsr, Sr = o.neumann_solve(oc, vnp)
    f, df, s, ds = ilm.neumann_poisson(cache, vnp, S=Sr)
    tol = 200 * np.linalg.cond(Sr) * np.finfo(float).eps
    assert relerr(df.data, dfr) < tol and relerr(ds.data, dsr) < tol
    assert relerr(f.array(), fr) < 1e-9 and relerr(s.array(), sr) < 1e-9
    f2, df2, _, _ = ilm.neumann_poisson(cache, vnp)          # S built on the GPU
    assert relerr(df2.data, dfr) < tol


def test_ifherk_stage_schur_complement(c1):
    """S_i = -E exp(L a) R, a = Fo/2 = 0.5: the integrating-factor
    kernel on the same engine, probed like create_RTLinvR."""
    cache, oc = c1
    g = cache.g
    E = ilm.lgf.intfact_table(0.5, g.NX)
    kid = cache.add_kernel(E)
    S = ilm.create_RTHR(cache, kid)
    plan = o.ConvPlan(E[:g.NX, :g.NY])
    Sref = np.zeros((cache.N, cache.N))
    for c in range(cache.N):
        e = np.zeros(cache.N); e[c] = 1.0
        Sref[:, c] = -oc.interpolate(plan.apply(oc.regularize(e)))
    assert relerr(S, Sref) < RTOL
    # a = 0: H = I, S_3 = -E R (stage 3 of LiskaIFHERK)
    kid0 = cache.add_kernel(ilm.lgf.intfact_table(0.0, g.NX))
    S3 = ilm.create_RTHR(cache, kid0)
    ER = -(o.E_matrix(oc.tabs[o.PRIMAL]) @ o.R_matrix(oc.tabs[o.PRIMAL])).toarray()
    assert relerr(S3, ER) < RTOL


def test_implicit_diffusion_kernel(c1):
    """implicit_operator(L, a) -> plan_implicit_diffusion(a L.factor) (src/grid_operators.jl:182-184):
    (I - a L)^-1 as a convolution on the same engine; (I - a L) applied to the result gives the input back."""
    cache, oc = c1
    g = cache.g
    a_tab = 0.5                                        # = a * L.factor
    K = ilm.lgf.implicit_diffusion_table(a_tab, g.NX)
    kid = cache.add_kernel(K)
    w = np.zeros(g.layout_shape(L.NODES_DUAL))
    w[40:90, 35:80] = np.random.default_rng(12).standard_normal((50, 45))
    u = ilm.Nodes(ilm.Dual, g).set(w)
    L.check(cache._lib.ilm_convolve(cache._plan, kid, L.NODES_DUAL, u.data.ctypes.data))
    assert relerr(u.array(), o.ConvPlan(K[:g.NX, :g.NY]).apply(w)) < RTOL
    lap = ilm.Nodes(ilm.Dual, g)
    ilm.laplacian(lap, u, cache)                       # scaled by L.factor = 1/dx^2
    back = u.array() - a_tab * g.dx ** 2 * lap.array()
    assert np.abs(back - w)[1:-1, 1:-1].max() < 1e-12 * np.abs(w).max()


# ---------------------------------------------------------------- convective terms (src/grid_operators.jl:258-434)
@pytest.mark.parametrize("case", ["c1", "odd"])
@pytest.mark.parametrize("device", [False, True])
def test_convective_terms(case, device, request):
    """The fused sweeps against the oracle's composition of grad! / grid_interpolate! / product! / transpose!
    (bit-exact: same operation order, no FMA contraction); smoke sequence of test/surface_ops.jl:258-285."""
    cache0, oc = request.getfixturevalue(case)
    g = cache0.g
    cache = (ilm.SurfaceScalarCache((oc.x, oc.y, oc.nx, oc.ny, oc.ds), g, lgf_table=ilm.lgf.lgf_table(max(g.NX, g.NY)), device=True)
             if device else cache0)
    rng = np.random.default_rng(23)

    def host(a):
        return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)

    def edges(u, v):
        return ilm.Edges(g, device=device).set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))

    us, vs = g.layout_shape(L.XEDGES), g.layout_shape(L.YEDGES)
    u, v = rng.standard_normal(us), rng.standard_normal(vs)
    u2, v2 = rng.standard_normal(us), rng.standard_normal(vs)
    p = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    w = rng.standard_normal(g.layout_shape(L.NODES_DUAL))
    q, q2 = edges(u, v), edges(u2, v2)
    dx = g.dx
    # v . grad p
    out = ilm.Nodes(ilm.Primal, g, device=device)
    ilm.convective_derivative(out, q, ilm.Nodes(ilm.Primal, g, device=device).set(p), cache)
    assert np.array_equal(host(out.data).reshape(p.shape, order="F"), o.convective_derivative_scalar(oc.grid, u, v, p, dx))
    # v . grad w on the dual nodes
    outd = ilm.Nodes(ilm.Dual, g, device=device)
    ilm.convective_derivative(outd, q, ilm.Nodes(ilm.Dual, g, device=device).set(w), cache)
    assert np.array_equal(host(outd.data).reshape(w.shape, order="F"), o.convective_derivative_dual(oc.grid, u, v, w, dx))
    # (v . grad) v and (v . grad) u
    nu = us[0] * us[1]
    for other, (ou_, ov_) in ((q, (u, v)), (q2, (u2, v2))):
        res = ilm.Edges(g, device=device)
        if other is q:
            ilm.convective_derivative(res, q, cache)
        else:
            ilm.convective_derivative(res, q, other, cache)
        ru, rv = o.convective_derivative_vector(oc.grid, u, v, ou_, ov_, dx)
        r = host(res.data)
        assert np.array_equal(r[:nu].reshape(us, order="F"), ru) and np.array_equal(r[nu:].reshape(vs, order="F"), rv)
    # w x v
    res = ilm.Edges(g, device=device)
    ilm.w_cross_v(res, ilm.Nodes(ilm.Dual, g, device=device).set(w), q, cache)
    ru, rv = o.w_cross_v(oc.grid, w, u, v)
    r = host(res.data)
    assert np.array_equal(r[:nu].reshape(us, order="F"), ru) and np.array_equal(r[nu:].reshape(vs, order="F"), rv)
    with pytest.raises(ilm.MethodError):
        ilm.convective_derivative(q, q, cache)


def test_convective_terms_full_size():
    """4096^2: (u . grad) u of a rigid rotation is the centripetal field -Omega^2 r, w x v likewise."""
    g = ilm.PhysicalGrid.centered(4096)
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g, device=True)
    xu, yu = g.coordinates(L.XEDGES)
    xv, yv = g.coordinates(L.YEDGES)
    Om = 0.75
    u = -Om * np.broadcast_to(yu[None, :], g.layout_shape(L.XEDGES))
    v = Om * np.broadcast_to(xv[:, None], g.layout_shape(L.YEDGES))
    q = ilm.Edges(g, device=True).set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    res = ilm.Edges(g, device=True)
    ilm.convective_derivative(res, q, cache)
    r = res.data.cpu().numpy()
    nu = u.size
    ru = r[:nu].reshape(u.shape, order="F")
    rv = r[nu:].reshape(v.shape, order="F")
    assert np.abs(ru[2:-2, 2:-2] + Om ** 2 * xu[2:-2, None]).max() < 1e-9
    assert np.abs(rv[2:-2, 2:-2] + Om ** 2 * yv[None, 2:-2]).max() < 1e-9
    w = ilm.Nodes(ilm.Dual, g, device=True).fill(2 * Om)              # vorticity of the rotation
    ilm.w_cross_v(res, w, q, cache)
    r = res.data.cpu().numpy()
    assert np.abs(r[:nu].reshape(u.shape, order="F")[2:-2, 2:-2] + 2 * Om ** 2 * xu[2:-2, None]).max() < 1e-9


# ---------------------------------------------------------------- band pass of the Schur probes (ilm_band.cu)
@pytest.mark.parametrize("ddf", ["yang3", "m4prime", "roma"])
def test_probe_band_pass_equals_transform_pass(ddf, monkeypatch):
    """The Schur probes replace the column transform by the direct sum over the patch rows
    (Y_n = sum_r x_r Gx(|n-r|), csrc/ilm_band.cu).  ILM_PROBE_BAND=0 keeps the transform column pass: both
    routes give the same matrices, for every builder, on an odd non-square grid with clipped windows,
    and with another kernel (integrating factor)."""
    g = ilm.PhysicalGrid(91, 67, 4.0 / 89, (45, 33))
    x, y, nx, ny, ds = ilm.bodies.circle(1.0, 1.4 * g.dx)
    body = (x + 0.93, y - 0.4, nx, ny, ds)                   # windows clipped at +x and -y
    G = ilm.lgf.lgf_table(96)
    built = {}
    for band in ("1", "0", "nopatch", "nosymm"):
        # "nopatch": band pass fed by pass A on the grid rows (ILM_PROBE_PATCH=0); default: the band pass sums the
        # x-spectrum of the DDF windows itself (create_RTLinvR probes: no pre-operator, no pass A) and create_RTLinvR
        # probes every column from its own window rows upwards only, the rest by symmetry ("nosymm": ILM_SCHUR_SYMM=0)
        monkeypatch.setenv("ILM_PROBE_BAND", "0" if band == "0" else "1")
        monkeypatch.setenv("ILM_PROBE_PATCH", "0" if band == "nopatch" else "1")
        monkeypatch.setenv("ILM_SCHUR_SYMM", "0" if band == "nosymm" else "1")
        cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, ddftype=ddf)
        vc = ilm.SurfaceVectorCache(body, g, lgf_table=G, ddftype=ddf)
        kid = cache.add_kernel(ilm.lgf.intfact_table(0.5, 96))
        built[band] = [ilm.create_RTLinvR(cache), ilm.create_CLinvCT(cache), ilm.create_GLinvD(cache),
                       ilm.create_GLinvD_cross(cache, cols=(3, 10)), ilm.create_RTHR(cache, kid),
                       ilm.create_CL2invCT(vc), ilm.create_CLinvCT(vc), ilm.create_RTLinvR(vc), ilm.create_GLinvD(vc, cols=(0, 9))]
    for a, b, c, d in zip(built["1"], built["0"], built["nopatch"], built["nosymm"]):
        assert relerr(a, b) < 1e-13
        assert relerr(a, c) < 1e-13
        assert relerr(a, d) < 1e-13
    if ddf == "yang3":
        oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body, G)
        assert relerr(built["1"][0], oc.create_RTLinvR()) < RTOL
        assert relerr(built["1"][0], oc.create_RTLinvR_table()) < RTOL


@pytest.mark.parametrize("cols", [None, (0, 7), (3, 40)])
def test_probe_pairs_off_the_band_path_inside_a_build(cols):
    """Two bodies far apart in y: the column pair that straddles the body boundary has patch rows more than
    16 apart and takes the transform column pass, the pairs around it take the band pass in patch mode with batched
    post sums.  The mixture (and odd column ranges) must give the same matrix as the oracle."""
    g = ilm.PhysicalGrid(120, 150, 4.0 / 118, (60, 75))
    c1_ = ilm.bodies.circle(0.45, 1.4 * g.dx, center=(0.3, -1.2))
    c2_ = ilm.bodies.circle(0.35, 1.4 * g.dx, center=(-0.5, 1.1))
    n1 = len(c1_[0])
    body = ilm.bodies.concat(c1_, c2_)[:5]
    G = ilm.lgf.lgf_table(160)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G)
    oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body, G)
    if cols is None:
        # make sure the boundary pair exists: start the build so that (n1 - 1, n1) is a pair
        lo = (n1 - 1) % 2
        S = np.asarray(ilm.create_RTLinvR(cache, cols=(lo, cache.N)))
        assert relerr(S, oc.create_RTLinvR()[:, lo:]) < RTOL
    else:
        S = np.asarray(ilm.create_RTLinvR(cache, cols=cols))
        assert relerr(S, oc.create_RTLinvR(cols=list(range(*cols)))) < RTOL


@pytest.mark.parametrize("scaling", [ilm.GridScaling, ilm.IndexScaling])
def test_symmetric_schur_build(scaling):
    """create_RTLinvR over all columns probes every pair from its own window rows upwards and mirrors the rest
    (S[k,c] / wgt_c is symmetric; wgt = ds/dx^2 with GridScaling, 1 with IndexScaling).  Two bodies with a gap in y (one
    pair falls back to full-row probes of single columns), windows clipped at the lower boundary, an odd number of points:
    the matrix equals the oracle's column-by-column probes and its FFT-free table form."""
    g = ilm.PhysicalGrid(120, 150, 4.0 / 118, (60, 75))
    c1_ = ilm.bodies.circle(0.45, 1.4 * g.dx, center=(0.3, -2.15))      # windows clipped at -y
    c2_ = ilm.bodies.circle(0.35, 1.4 * g.dx, center=(-0.5, 1.1))
    body = ilm.bodies.concat(c1_, c2_)[:5]
    if len(body[0]) % 2 == 0:
        body = tuple(a[:-1] for a in body)
    G = ilm.lgf.lgf_table(160)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, scaling=scaling)
    oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body, G, scaling=o.GRID_SCALING if scaling == ilm.GridScaling else o.INDEX_SCALING)
    S = np.asarray(ilm.create_RTLinvR(cache))
    assert S.shape == (cache.N, cache.N) and cache.N % 2 == 1
    assert relerr(S, oc.create_RTLinvR()) < RTOL
    assert relerr(S, oc.create_RTLinvR_table()) < RTOL
    E = ilm.lgf.intfact_table(0.7, 160)                                # any even kernel: stage complements of IF-HERK
    kid = cache.add_kernel(E)
    plan = o.ConvPlan(E[:g.NX, :g.NY])
    cols = [0, 1, cache.N // 2, cache.N - 1]
    Sref = np.zeros((cache.N, len(cols)))
    for q, c in enumerate(cols):
        e = np.zeros(cache.N); e[c] = 1.0
        Sref[:, q] = -oc.interpolate(plan.apply(oc.regularize(e)))
    assert relerr(np.asarray(ilm.create_RTHR(cache, kid))[:, cols], Sref) < RTOL


@pytest.mark.parametrize("ddf", ["yang3", "m4prime"])
def test_device_list_build_equals_host_build(ddf, monkeypatch):
    """The gather lists (cell buckets, row buckets, column mask of pass C) are built by a radix sort on the device;
    ILM_TABLES_HOST=1 keeps the host build.  Same lists -> bit-equal regularization fields and Schur complements,
    also for windows clipped at the boundary and after a refresh with moved points."""
    g = ilm.PhysicalGrid(91, 67, 4.0 / 89, (45, 33))
    x, y, nx, ny, ds = ilm.bodies.circle(1.0, 1.4 * g.dx)
    body = (x + 0.93, y - 0.4, nx, ny, ds)
    G = ilm.lgf.lgf_table(96)
    rng = np.random.default_rng(3)
    f = rng.standard_normal(len(x))
    out = {}
    for mode in ("device", "host"):
        if mode == "host":
            monkeypatch.setenv("ILM_TABLES_HOST", "1")
        else:
            monkeypatch.delenv("ILM_TABLES_HOST", raising=False)
        cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, ddftype=ddf)
        fs = cache.zeros_surface(); fs.data[:] = f
        w = cache.zeros_grid(); ilm.regularize(w, fs, cache)
        q = cache.zeros_gridgrad(); ilm.regularize_normal(q, fs, cache)
        S = np.asarray(ilm.create_RTLinvR(cache))
        Cf = np.asarray(ilm.create_surface_filter(cache))
        cache.update_points((x + 0.5, y - 0.1, nx, ny, ds))
        w2 = cache.zeros_grid(); ilm.regularize(w2, fs, cache)
        S2 = np.asarray(ilm.create_RTLinvR(cache, cols=(0, 11)))
        out[mode] = [np.asarray(w.data).copy(), np.asarray(q.data).copy(), S, Cf, np.asarray(w2.data).copy(), S2]
    for a, b in zip(out["device"], out["host"]):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------- whole-problem entry point (ilm_dirichlet_poisson)
@pytest.mark.parametrize("device", [False, True])
def test_dirichlet_solve_single_call(device):
    """`solve(prob, sys)` of test/literate/dirichlet.jl:71-107 as one library call: same S (bit-equal: the same
    kernels), field and multiplier as the operator-by-operator composition, and as the oracle."""
    g = ilm.PhysicalGrid.centered(128)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(128)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=device)
    oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G)
    fplus, fminus = cache.points()[0].copy(), 0.3 * cache.points()[1]
    f, s, S = ilm.dirichlet_solve(cache, fplus, fminus, return_S=True)
    f1, s1, S1 = ilm.dirichlet_poisson(cache, fplus, fminus)
    tonp = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)   # noqa: E731
    assert np.array_equal(tonp(S), tonp(S1))
    assert relerr(tonp(f.data), tonp(f1.data)) < 1e-13 and relerr(tonp(s.data), tonp(s1.data)) < 1e-9
    fr, sr, Sr = o.dirichlet_solve(oc, fplus, fminus)
    assert relerr(tonp(S), Sr) < RTOL
    assert relerr(f.array(), fr) < 1e-10
    assert relerr(tonp(s.data), sr) < 50 * np.linalg.cond(Sr) * np.finfo(float).eps
    f2, s2 = ilm.dirichlet_solve(cache, fplus)                   # fminus = None
    fr2, _, _ = o.dirichlet_solve(oc, fplus, S=Sr)
    assert relerr(f2.array(), fr2) < 1e-10
    f3, s3 = ilm.dirichlet_solve(cache, fplus, want_field=False)  # a rank that only wants the multiplier (f = NULL)
    assert f3 is None and np.array_equal(tonp(s3.data), tonp(s2.data))
    mx, my = g.layout_shape(L.NODES_PRIMAL)
    f4, s4 = ilm.dirichlet_solve(cache, fplus, field_rows=(17, 90))   # a rank that keeps a slab of rows (ilm_dirichlet_poisson_rows)
    assert np.array_equal(tonp(f4), tonp(f2.data)[17 * mx:90 * mx]) and np.array_equal(tonp(s4.data), tonp(s2.data))
    with pytest.raises(ilm.DimensionMismatch):
        ilm.dirichlet_solve(cache, fplus, field_rows=(90, my + 1))
    assert cache.comm_info() == (0, 1)
    assert relerr(tonp(ilm.create_schur_sharded(cache, "CLinvCT")), oc.create_CLinvCT()) < RTOL


def test_goza_ddf_is_named_and_refused():
    """src/cache.jl:305 lists Goza among the DDF types; its kernel exists only in the un-vendored CartesianGrids, so the
    library names it and answers ILM_EINVAL rather than guessing a formula."""
    g = ilm.PhysicalGrid.centered(32)
    with pytest.raises(ilm.MethodError, match="Goza"):
        ilm.SurfaceScalarCache(ilm.bodies.circle(0.5, 1.4 * g.dx), g, ddftype="goza")


@pytest.mark.parametrize("device", [False, True])
def test_complex_cache(device):
    """`SurfaceScalarCache(body, g, dtype=ComplexF64)` (src/cache.jl:167,187; test/surface_ops.jl:109-119): the
    similar_* containers are complex, and the (real-linear) operators act on real and imaginary parts."""
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(64)
    cc = ilm.SurfaceScalarCache(body, g, lgf_table=G, dtype=complex, device=device)
    rc = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=device)
    for make in (cc.similar_grid, cc.similar_gridcurl, cc.similar_gridgrad, cc.similar_surface):
        assert make().eltype is complex                       # test/surface_ops.jl:112-116
    assert rc.similar_grid().eltype is float
    vc = ilm.SurfaceVectorCache(body, g, lgf_table=G, dtype=complex, device=device)
    assert vc.similar_grid().eltype is complex and vc.similar_surface().eltype is complex
    rng = np.random.default_rng(4)
    fr, fi = rng.standard_normal(cc.N), rng.standard_normal(cc.N)
    f = cc.similar_surface().set(fr + 1j * fi)
    s = cc.similar_grid()
    ilm.regularize(s, f, cc)
    ilm.inverse_laplacian(s, cc)
    out = cc.similar_surface()
    ilm.interpolate(out, s, cc)
    ref = []
    for part in (fr, fi):
        sp = rc.zeros_grid()
        ilm.regularize(sp, rc.zeros_surface().set(part), rc)
        ilm.inverse_laplacian(sp, rc)
        o_ = rc.zeros_surface()
        ilm.interpolate(o_, sp, rc)
        ref.append(o_.numpy().copy())
    assert np.array_equal(out.numpy(), ref[0] + 1j * ref[1])
    with pytest.raises(ilm.MethodError):
        ilm.regularize(rc.zeros_grid(), f, cc)                # real and complex containers do not mix
    with pytest.raises(ilm.MethodError):
        ilm.convective_derivative(cc.similar_gridgrad(), cc.similar_gridgrad(), cc)


def test_far_field_constant_is_the_callers_parameter():
    """c0 of `L\\w` is not pinned by any reference test: it is an explicit argument of ilm_plan_create and
    the library uses exactly the caller's value -- two plans that differ only in c0 differ by -dc0 sum(w) / factor in
    L^-1 w, and by the matching rank-one term -dc0/factor (E 1)(1^T R) in S = -E L^-1 R."""
    g = ilm.PhysicalGrid.centered(64)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(64)
    ca = ilm.SurfaceScalarCache(body, g, lgf_table=G)                     # default: (gamma + ln(8)/2 - ln dx) / 2 pi
    cb = ilm.SurfaceScalarCache(body, g, lgf_table=G, c0=0.25)
    assert abs(ca.c0 - ilm.lgf.lgf_c0(g.dx)) < 1e-15 and cb.c0 == 0.25
    w = np.random.default_rng(9).standard_normal(g.layout_shape(L.NODES_PRIMAL))
    wa, wb = ca.zeros_grid().set(w), cb.zeros_grid().set(w)
    ilm.inverse_laplacian(wa, ca)
    ilm.inverse_laplacian(wb, cb)
    shift = -(ca.c0 - cb.c0) * w.sum() / ca.lap_factor
    assert np.abs((wa.array() - wb.array()) - shift).max() < 1e-12 * max(np.abs(wa.array()).max(), abs(shift))
    Sa, Sb = ilm.create_RTLinvR(ca), ilm.create_RTLinvR(cb)
    _, wR, wE = ca.table(L.NODES_PRIMAL)
    rank1 = (ca.c0 - cb.c0) / ca.lap_factor * np.outer(wE.sum(axis=(1, 2)), wR.sum(axis=(1, 2)))
    assert np.abs((Sa - Sb) - rank1).max() < 1e-12 * np.abs(Sa).max()

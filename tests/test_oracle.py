"""Pin the CPU oracle against everything the reference holds for the hot path
(SURVEY.md section 4 identity table and Appendix B golden values).  CPU only."""
import numpy as np
import pytest

import ilm_b200
import ilm_oracle as o

bodies = ilm_b200.bodies
lgfmod = ilm_b200.lgf


def layers_setup():
    """examples/Layers.ipynb cells 5,19: 600^2 grid, dx=0.02, circle R=1, ds=1.5dx."""
    g = o.Grid(600, 600, 0.02, (300, 300))
    x, y, nx, ny, ds = bodies.circle(1.0, 1.5 * 0.02)
    return g, x, y, nx, ny, ds


# ---------------------------------------------------------------- DDFs
@pytest.mark.parametrize("name", list(o.DDFS))
def test_ddf_moments(name):
    """Partition of unity and first moment (SURVEY.md A.2 verification)."""
    fn, rho, W = o.DDFS[name]
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, 200)
    j = np.arange(-4, 6)
    phi = fn(x[:, None] - j[None, :])
    assert np.abs(phi.sum(axis=1) - 1).max() < 4e-15
    if name != "witchhat":
        assert np.abs(((x[:, None] - j[None, :]) * phi).sum(axis=1)).max() < 4e-15


def test_yang3_center_value():
    """phi(0) of Yang3."""
    assert abs(o.ddf_yang3(np.array([0.0]))[0] - 0.6181999293593575) < 2e-16


def test_asin_fdlibm_matches_libm():
    x = np.concatenate([np.linspace(-1, 1, 20001), np.random.default_rng(1).uniform(-1, 1, 20000)])
    ref = np.arcsin(x)
    got = o.asin_fdlibm(x)
    ulp = np.spacing(np.abs(ref)) + 1e-300
    assert np.max(np.abs(got - ref) / ulp) <= 1.0


# ---------------------------------------------------------------- golden values
def test_golden_circle_sum_ds():
    """examples/caches.ipynb:1380: dot(1,1,cache) = 6.283288300933757 for
    Circle(1.0, 0.014) (448 points); our midpoint-rule circle reproduces it to
    4e-12 (Appendix B)."""
    x, y, nx, ny, ds = bodies.circle(1.0, 1.4 * 0.01)
    assert x.shape[0] == 448
    assert abs(ds.sum() - 6.283288300933757) < 1e-11
    assert abs(nx[1] - 0.9999016728287639) < 1e-7      # caches.ipynb:1177
    x2 = bodies.circle(0.5, 0.014)[0]
    assert x2.shape[0] == 224                             # multbodies.ipynb:77


def test_golden_layers_dot_x_Rf_n():
    """examples/Layers.ipynb cell 64: dot(qx, Rf*nrm, g) = 3.141119452036813 with
    regop weights = dlengthmid(body) (cell 19) and Yang3 on the u-edges."""
    g, x, y, nx, ny, ds = layers_setup()
    N = x.shape[0]
    assert N == 209                                       # cell 62 output
    dlmid = np.full(N, np.sin(2 * np.pi / N))            # 0.5*|x[k+1]-x[k-1]|
    tab = o.build_table(g, x, y, dlmid, o.XEDGE, "yang3", o.GRID_SCALING)
    q = o.regularize(tab, nx)
    xc, _ = g.coords(o.XEDGE)
    qx = np.broadcast_to(xc[:, None], q.shape)
    assert abs(o.dot_grid(g, qx, q, o.XEDGE) - 3.141119452036813) < 2e-14


def test_golden_layers_grid_coordinates():
    """Layers.ipynb cells 54-55: v-edge coords -5.98:0.02:5.98 x -5.99:0.02:5.99, xv[300]=0."""
    g = o.Grid(600, 600, 0.02, (300, 300))
    xv, yv = g.coords(o.YEDGE)
    assert xv.shape[0] == 599 and yv.shape[0] == 600
    assert abs(xv[0] + 5.98) < 1e-12 and abs(yv[0] + 5.99) < 1e-12 and xv[299] == 0.0


def test_golden_integrate_ones():
    """caches.ipynb:1432: integrate(ones) on the 406^2 grid = 16.3216."""
    g = o.Grid(406, 406, 0.01, (203, 203))
    one = np.ones(o.field_shape(o.PRIMAL, 406, 406))
    assert abs(o.dot_grid(g, one, one, o.PRIMAL) - 16.3216) < 1e-10


# ---------------------------------------------------------------- reference test identities
def test_tools_jl_regularization_identities():
    """test/tools.jl:96-122 on its fixture (600^2, dx=0.02, Circle(1,1.5dx))."""
    g, x, y, nx, ny, ds = layers_setup()
    N = x.shape[0]
    tab = o.build_table(g, x, y, ds, o.DUAL, "yang3", o.GRID_SCALING)
    oc = np.ones(o.field_shape(o.DUAL, 600, 600))
    os_ = np.ones(N)
    # :107  dot(oc, Rc*os, g) == dot(os, os, ds)
    assert abs(o.dot_grid(g, oc, o.regularize(tab, os_), o.DUAL) - o.dot_surface(os_, os_, ds)) < 1e-14 * 10
    # :112 adjointness with random data
    rng = np.random.default_rng(0)
    u = rng.standard_normal(oc.shape)
    u[0, :] = u[-1, :] = 0
    u[:, 0] = u[:, -1] = 0
    phi = rng.standard_normal(N)
    lhs = o.dot_grid(g, u, o.regularize(tab, phi), o.DUAL)
    rhs = o.dot_surface(o.interpolate(tab, u), phi, ds)
    assert abs(lhs - rhs) < 1e-13
    # :119-120 Rf on the u component
    tu = o.build_table(g, x, y, ds, o.XEDGE, "yang3", o.GRID_SCALING)
    qu = np.ones(o.field_shape(o.XEDGE, 600, 600))
    val = o.dot_grid(g, qu, o.regularize(tu, os_), o.XEDGE)
    assert abs(val - ds.sum()) < 1e-12
    assert abs(val - 2 * np.pi) < 1e-3


def test_regularize_matches_csc_matvec():
    """The add.at restatement equals scipy's CSC mat-vec (the structure upstream uses)."""
    g = o.Grid(64, 48, 0.05, (32, 24))
    x, y, nx, ny, ds = bodies.circle(0.7, 1.4 * 0.05)
    rng = np.random.default_rng(3)
    f = rng.standard_normal(x.shape[0])
    for kind in o.KINDS:
        tab = o.build_table(g, x, y, ds, kind)
        a = o.regularize(tab, f)
        b = (o.R_matrix(tab) @ f).reshape(tab.shape, order="F")
        assert np.array_equal(a, b)
        s = rng.standard_normal(tab.shape)
        fa = o.interpolate(tab, s)
        fb = o.E_matrix(tab) @ s.ravel(order="F")
        assert np.abs(fa - fb).max() < 1e-14


def test_index_scaling_symmetric():
    g = o.Grid(64, 64, 0.05, (32, 32))
    x, y, nx, ny, ds = bodies.circle(0.7, 0.07)
    tab = o.build_table(g, x, y, ds, o.PRIMAL, "yang3", o.INDEX_SCALING)
    assert np.array_equal(tab.wR, tab.wE)


# ---------------------------------------------------------------- stencil identities
def test_stencil_identities():
    """D=-G^T, C^T=transpose(C), DC=0, C^T G=0."""
    g = o.Grid(20, 17, 1.0, (10, 8))
    rng = np.random.default_rng(5)
    NX, NY = g.NX, g.NY
    p = rng.standard_normal(o.field_shape(o.PRIMAL, NX, NY))
    s = rng.standard_normal(o.field_shape(o.DUAL, NX, NY))
    u = rng.standard_normal(o.field_shape(o.XEDGE, NX, NY))
    v = rng.standard_normal(o.field_shape(o.YEDGE, NX, NY))
    # interior-supported data so boundary truncation plays no role
    for a in (p, s, u, v):
        a[:2, :] = 0; a[-2:, :] = 0; a[:, :2] = 0; a[:, -2:] = 0
    gu, gv = o.grad_n2e(g, p)
    lhs = np.sum(gu * u) + np.sum(gv * v)
    rhs = -np.sum(p * o.divergence_e2n(g, u, v))
    assert abs(lhs - rhs) < 1e-12
    cu, cv = o.curl_n2e(g, s)
    assert abs(np.sum(cu * u) + np.sum(cv * v) - np.sum(s * o.curl_e2n(g, u, v))) < 1e-12
    assert np.abs(o.divergence_e2n(g, cu, cv)).max() < 1e-13
    assert np.abs(o.curl_e2n(g, gu, gv)).max() < 1e-13
    # L_F = G D - C C^T on edges (test/literate/matrices.jl:40-45)
    gdu, gdv = o.grad_n2e(g, o.divergence_e2n(g, u, v))
    ccu, ccv = o.curl_n2e(g, o.curl_e2n(g, u, v))
    lu = o.laplacian(g, u, o.XEDGE)
    assert np.abs((gdu - ccu) - lu)[3:-3, 3:-3].max() < 1e-12
    # tensor pair: D_t = -G_t^T
    t = o.grad_e2t(g, u, v)
    tt = [rng.standard_normal(a.shape) for a in t]
    for a in tt:
        a[:2, :] = 0; a[-2:, :] = 0; a[:, :2] = 0; a[:, -2:] = 0
    du, dv = o.divergence_t2e(g, *tt)
    assert abs(sum(np.sum(a * b) for a, b in zip(t, tt)) + np.sum(du * u) + np.sum(dv * v)) < 1e-12


# ---------------------------------------------------------------- LGF
def test_lgf_exact_values_and_delta():
    G = lgfmod.lgf_table(300)
    assert abs(G[1, 0] - 0.25) < 1e-15
    assert abs(G[1, 1] - 1 / np.pi) < 1e-15
    assert abs(G[2, 0] - (1 - 2 / np.pi)) < 1e-15
    L = G[2:, 1:-1] + G[:-2, 1:-1] + G[1:-1, 2:] + G[1:-1, :-2] - 4 * G[1:-1, 1:-1]
    assert np.abs(L).max() < 5e-14
    assert abs(4 * G[1, 0] - 4 * G[0, 0] - 1) < 1e-15        # L G = +delta at the origin
    n = np.arange(1, 300)
    assert np.abs(np.diag(G)[1:] - np.cumsum(1 / (2 * n - 1)) / np.pi).max() < 5e-15
    Gr = lgfmod.lgf_table(120, rule="gl100")
    assert np.abs(Gr - G[:120, :120]).max() < 1e-12


def test_fft_convolution_matches_direct():
    G = lgfmod.lgf_table(40)
    rng = np.random.default_rng(2)
    plan = o.ConvPlan(G[:24, :20])
    for shape in [(24, 20), (23, 19), (24, 19), (23, 20)]:
        w = rng.standard_normal(shape)
        a = plan.apply(w)
        b = o.direct_convolution(G, w)
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()


def test_inverse_laplacian_inverts_laplacian():
    """L (L^-1 w) = w for compactly supported w away from the boundary."""
    g = o.Grid(64, 64, 0.1, (32, 32))
    G = lgfmod.lgf_table(64)
    plan = o.ConvPlan(G)
    w = np.zeros(o.field_shape(o.DUAL, 64, 64), order="F")
    w[20:40, 22:41] = np.random.default_rng(0).standard_normal((20, 19))
    factor = 1 / g.dx ** 2
    sol = o.inverse_laplacian(plan, w, o.lgf_c0(g.dx), factor)
    back = o.laplacian(g, sol, o.DUAL, factor)
    assert np.abs(back - w)[1:-1, 1:-1].max() < 1e-11 * np.abs(w).max()


# ---------------------------------------------------------------- physics checks (test/surface_ops.jl)
@pytest.fixture(scope="module")
def small_cache():
    """test/surface_ops.jl:4-22 fixture scaled down: 4x4 domain dx=0.04, R=1, ds=1.4dx."""
    dx = 0.04
    NX = 104
    g = o.Grid(NX, NX, dx, (NX // 2, NX // 2))
    x, y, nx, ny, ds = bodies.circle(1.0, 1.4 * dx)
    G = lgfmod.lgf_table(NX)
    return o.ScalarCache(g, x, y, nx, ny, ds, G)


def test_mask_integral(small_cache):
    """test/surface_ops.jl:180-186: integral of the mask = pi R^2 (2e-3)."""
    c = small_cache
    m = c.mask()
    area = o.dot_grid(c.grid, m, np.ones_like(m), o.PRIMAL)
    assert abs(area - np.pi) < 2e-2        # coarser grid than the reference fixture
    assert abs(m[51, 51] - 1.0) < 1e-3 and abs(m[2, 2]) < 1e-3


def test_schur_operator_norms(small_cache):
    """test/surface_ops.jl:63-77: nRTRn max singular value ~ 11, CLinvCT(scale=dx)
    max eigenvalue ~ 0.2, GLinvD max eigenvalue ~ 0.45."""
    c = small_cache
    A = c.create_nRTRn()
    assert abs(np.linalg.svd(A, compute_uv=False).max() - 11) < 1.5
    cols = range(0, c.N, 1)
    Cm = c.create_CLinvCT(scale=c.grid.dx, cols=cols)
    assert abs(np.abs(np.linalg.eigvals(Cm)).max() - 0.2) < 0.1
    Gm = c.create_GLinvD(scale=c.grid.dx, cols=cols)
    th = 2 * np.pi * np.arange(c.N) / c.N
    fn = c.normal_interpolate(*c.regularize_normal(np.sin(th - np.pi / 4)))
    assert abs(fn.max() - 11) < 1.5 and abs(fn.min() + 11) < 1.5
    assert abs(np.abs(np.linalg.eigvals(Gm)).max() - 0.45) < 0.1


def test_dirichlet_solution(small_cache):
    """test/literate/dirichlet.jl: phi+ = x outside, 0 inside."""
    c = small_cache
    f, s, S = o.dirichlet_solve(c, c.x.copy())
    # interior is blank, exterior matches the analytic exterior solution x/r^2 near the body
    assert np.abs(f[51, 51]) < 5e-3
    xg, yg = c.grid.coords(o.PRIMAL)
    i, j = 51 + 35, 51          # (1.4, 0)
    assert abs(f[i, j] - xg[i] / (xg[i] ** 2 + yg[j] ** 2)) < 3e-2
    C = c.create_surface_filter()
    assert np.abs(C.sum(axis=1) - 1).max() < 1e-12       # filter preserves constants


def test_zero_body_cache():
    """test/surface_ops.jl:356-371: N = 0 caches work and give zeros."""
    g = o.Grid(32, 32, 0.1, (16, 16))
    z = np.zeros(0)
    c = o.ScalarCache(g, z, z, z, z, z, lgfmod.lgf_table(32))
    assert np.all(c.regularize(z) == 0)
    assert c.interpolate(g.zeros(o.PRIMAL)).shape == (0,)
    assert np.all(c.mask() == 1.0)
    assert c.create_RTLinvR().shape == (0, 0)


def test_c_direct_convolution_matches_fft_oracle():
    """oracle/direct_conv.c (long-double O(P^2)) vs the FFT restatement, incl. c0 and factor."""
    import ctypes
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    so = os.path.join(here, "libilm_oracle_c.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", here])
    lib = ctypes.CDLL(so)
    G = np.asfortranarray(lgfmod.lgf_table(40))
    plan = o.ConvPlan(G[:33, :29])
    rng = np.random.default_rng(11)
    dp = ctypes.c_void_p
    for shape in [(33, 29), (32, 28), (33, 28)]:
        w = np.asfortranarray(rng.standard_normal(shape))
        out = np.zeros(shape, order="F")
        c0, factor = o.lgf_c0(0.05), 400.0
        lib.ilm_oracle_inverse_laplacian(dp(G.ctypes.data), 40, dp(w.ctypes.data), shape[0], shape[1],
                                         ctypes.c_double(c0), ctypes.c_double(factor), dp(out.ctypes.data))
        ref = o.inverse_laplacian(plan, w, c0, factor)
        assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()


# ---------------------------------------------------------------- vector cache (test/surface_ops.jl:121-146)
@pytest.fixture(scope="module")
def small_vcache():
    dx = 0.04
    NX = 104
    g = o.Grid(NX, NX, dx, (NX // 2, NX // 2))
    x, y, nx, ny, ds = bodies.circle(1.0, 1.4 * dx)
    return o.VectorCache(g, x, y, nx, ny, ds, lgfmod.lgf_table(NX))


def test_vector_surface_ops_reference_values(small_vcache):
    """regularize_normal_symm! / normal_interpolate_symm! round trip: extrema of vs.u ~ +-22.5
    (test/surface_ops.jl:124-137, same sequence of calls)."""
    c = small_vcache
    th = 2 * np.pi * np.arange(c.N) / c.N
    vu, vv = np.sin(th - np.pi / 4), np.zeros(c.N)
    vu, vv = c.normal_interpolate_v(c.regularize_normal_v(vu, vv))
    vu = np.sin(th - np.pi / 4)
    vu, vv = c.normal_interpolate_symm(c.regularize_normal_symm(vu, vv))
    assert abs(vu.max() - 22.5) < 1.0 and abs(vu.min() + 22.5) < 1.0


def test_vector_schur_and_adjoints(small_vcache):
    c = small_vcache
    g = c.grid
    rng = np.random.default_rng(21)
    vu, vv = rng.standard_normal(c.N), rng.standard_normal(c.N)
    A = [rng.standard_normal(o.field_shape(k, g.NX, g.NY)) for k in (o.PRIMAL, o.DUAL, o.DUAL, o.PRIMAL)]
    for a in A:
        a[:2, :] = 0; a[-2:, :] = 0; a[:, :2] = 0; a[:, -2:] = 0
    # <A, Rt (n o v)>_grid dx^2 = <n . Et A, v>_ds   (regularize_normal! / normal_interpolate! adjoint pair)
    T = c.regularize_normal_v(vu, vv)
    lhs = sum(np.sum(a * t) for a, t in zip(A, T)) * g.dx ** 2
    tu, tv = c.normal_interpolate_v(A)
    assert abs(lhs - np.sum((tu * vu + tv * vv) * c.ds)) < 1e-10 * abs(lhs)
    Ts = c.regularize_normal_symm(vu, vv)
    lhs = sum(np.sum(a * t) for a, t in zip(A, Ts)) * g.dx ** 2
    tu, tv = c.normal_interpolate_symm(A)
    assert abs(lhs - np.sum((tu * vu + tv * vv) * c.ds)) < 1e-10 * abs(lhs)
    # surface_grad is the negative adjoint of surface_divergence (src/surface_operators.jl:557)
    qu = rng.standard_normal(o.field_shape(o.XEDGE, g.NX, g.NY)); qv = rng.standard_normal(o.field_shape(o.YEDGE, g.NX, g.NY))
    for a in (qu, qv):
        a[:3, :] = 0; a[-3:, :] = 0; a[:, :3] = 0; a[:, -3:] = 0
    du, dv = c.surface_divergence_v(vu, vv)
    gu, gv = c.surface_grad_v(qu, qv)
    lhs = (np.sum(du * qu) + np.sum(dv * qv)) * g.dx ** 2
    assert abs(lhs + np.sum((gu * vu + gv * vv) * c.ds)) < 1e-10 * abs(lhs)
    # create_GLinvD(vcache, scale=dx): max |eig| ~ 0.45 (test/surface_ops.jl:145-146), sampled columns keep it quick
    cols = list(range(0, 2 * c.N, 1))
    Gm = c.create_GLinvD_v(scale=g.dx, cols=cols)
    assert abs(np.abs(np.linalg.eigvals(Gm)).max() - 0.45) < 0.1
    # mask on Edges integrates to the body area
    mu, mv = c.mask_edges()
    assert abs(o.dot_grid(g, mu, np.ones_like(mu), o.XEDGE) - np.pi) < 3e-2


def test_neumann_added_mass(small_cache):
    """test/literate/neumann.jl: v_n+ = n_x on the unit circle; the potential jump df ~ -2 x and the
    added mass  -int df n_x ds  -> pi (examples/neumann.ipynb reports the coefficient of a small circle)."""
    c = small_cache
    f, df, s, ds, S = o.neumann_solve(c, c.nx.copy())
    assert abs(np.sum(df * c.nx * c.ds) + np.pi) < 0.15


def test_implicit_diffusion_table():
    """(I - aL) K = delta on the lattice and sum K = 1 (the symbol at k = 0)."""
    for a in (0.05, 0.5, 2.0):
        K = lgfmod.implicit_diffusion_table(a, 96)
        LK = K[2:, 1:-1] + K[:-2, 1:-1] + K[1:-1, 2:] + K[1:-1, :-2] - 4 * K[1:-1, 1:-1]
        assert np.abs(K[1:-1, 1:-1] - a * LK).max() < 1e-15
        assert abs(K[0, 0] - a * (2 * K[1, 0] + 2 * K[0, 1] - 4 * K[0, 0]) - 1) < 1e-15
        tot = 4 * K.sum() - 2 * K[0, :].sum() - 2 * K[:, 0].sum() + K[0, 0]
        assert abs(tot - 1) < 1e-13


# ---------------------------------------------------------------- mask products (test/surface_ops.jl:190-232)
def test_mask_products_integrate_to_area(small_cache, small_vcache):
    """mask!(ones) on every layout integrates to pi R^2 (the reference asserts 1e-3 on its dx=0.02 grid;
    this fixture is coarser), and mask + complementary mask reproduce the field away from the ghosts."""
    c, g = small_cache, small_cache.grid
    for kind in (o.PRIMAL, o.DUAL, o.XEDGE, o.YEDGE):
        one = np.ones(o.field_shape(kind, g.NX, g.NY))
        inner = c.mask_product(one, kind)
        assert abs(o.dot_grid(g, inner, one, kind) - np.pi) < 2e-2
        outer = c.mask_product(one, kind, complementary=True)
        assert np.abs((inner + outer)[2:-2, 2:-2] - 1.0).max() < 1e-12
    v = small_vcache
    ones_e = tuple(np.ones(o.field_shape(k, g.NX, g.NY)) for k in (o.XEDGE, o.YEDGE))
    mu, mv = v.mask_product_v(ones_e, "edges")
    assert abs(o.dot_grid(g, mu, ones_e[0], o.XEDGE) - np.pi) < 2e-2
    assert abs(o.dot_grid(g, mv, ones_e[1], o.YEDGE) - np.pi) < 2e-2
    kinds = (o.PRIMAL, o.DUAL, o.DUAL, o.PRIMAL)
    ones_t = tuple(np.ones(o.field_shape(k, g.NX, g.NY)) for k in kinds)
    for comp, k, one in zip(v.mask_product_v(ones_t, "edgegrad"), kinds, ones_t):
        assert abs(o.dot_grid(g, comp, one, k) - np.pi) < 2e-2


def test_grid_interpolate_is_exact_for_linear_fields():
    g = o.Grid(12, 9, 0.5, (3, 2))
    for src in o.KINDS:
        xs, ys = g.coords(src)
        m = 2.0 * xs[:, None] - 3.0 * ys[None, :] + 0.25
        for dst in o.KINDS:
            xd, yd = g.coords(dst)
            ref = 2.0 * xd[:, None] - 3.0 * yd[None, :] + 0.25
            got = o.grid_interpolate(g, m, src, dst)
            assert np.abs(got - ref)[2:-2, 2:-2].max() < 1e-13      # the outer ring may be left at 0


def test_convective_terms_on_polynomial_fields():
    """v . grad p, (v . grad) v and w x v of the oracle (src/grid_operators.jl:258-434 compositions) reproduce
    the continuous operators on fields for which the second-order staggered scheme is exact."""
    g = o.Grid(20, 16, 0.25, (5, 4))
    xu, yu = g.coords(o.XEDGE)
    xv, yv = g.coords(o.YEDGE)
    xp, yp = g.coords(o.PRIMAL)
    xd, yd = g.coords(o.DUAL)
    # uniform velocity (a, b), linear p: v . grad p = a px + b py exactly
    a, b, px, py = 0.7, -0.3, 2.0, 1.5
    u = np.full(o.field_shape(o.XEDGE, g.NX, g.NY), a)
    v = np.full(o.field_shape(o.YEDGE, g.NX, g.NY), b)
    p = px * xp[:, None] + py * yp[None, :]
    out = o.convective_derivative_scalar(g, u, v, p, div=g.dx)
    assert np.abs(out[2:-2, 2:-2] - (a * px + b * py)).max() < 1e-12
    # linear velocity field (u, v) = (x + 2y, 3x - y): (v . grad) v is linear, the scheme is exact
    u = xu[:, None] + 2 * yu[None, :]
    v = 3 * xv[:, None] - yv[None, :]
    ou, ov = o.convective_derivative_vector(g, u, v, u, v, div=g.dx)
    ru = (xu[:, None] + 2 * yu[None, :]) * 1.0 + (3 * xu[:, None] - yu[None, :]) * 2.0
    rv = (xv[:, None] + 2 * yv[None, :]) * 3.0 + (3 * xv[:, None] - yv[None, :]) * (-1.0)
    assert np.abs(ou[2:-2, 2:-2] - ru[2:-2, 2:-2]).max() < 1e-12
    assert np.abs(ov[2:-2, 2:-2] - rv[2:-2, 2:-2]).max() < 1e-12
    # w x v with uniform w: (w e_z) x (u, v) = (-w v, w u)
    w = np.full(o.field_shape(o.DUAL, g.NX, g.NY), 1.3)
    ou, ov = o.w_cross_v(g, w, u, v)
    assert np.abs(ou[2:-2, 2:-2] - (-1.3 * (3 * xu[:, None] - yu[None, :]))[2:-2, 2:-2]).max() < 1e-12
    assert np.abs(ov[2:-2, 2:-2] - (1.3 * (xv[:, None] + 2 * yv[None, :]))[2:-2, 2:-2]).max() < 1e-12
    # v . grad w on the dual nodes: uniform velocity, linear w
    u = np.full(o.field_shape(o.XEDGE, g.NX, g.NY), a)
    v = np.full(o.field_shape(o.YEDGE, g.NX, g.NY), b)
    wd = px * xd[:, None] + py * yd[None, :]
    out = o.convective_derivative_dual(g, u, v, wd, div=g.dx)
    assert np.abs(out[2:-2, 2:-2] - (a * px + b * py)).max() < 1e-12


# ---------------------------------------------------------------- forcing regions (src/forcing.jl, test/surface_ops.jl:445-542)
def test_forcing_region_restatement(small_cache):
    c = small_cache
    g = c.grid
    shp = o.field_shape(o.PRIMAL, g.NX, g.NY)
    # point forcing with M4': partition of unity -> sum(dT) dx^2 = sum(str); weights are 1/dx^2
    px, py = np.array([-1.2, 0.5, 0.013]), np.array([0.5, 0.5, -0.777])
    strength = np.array([1.0, -1.0, 2.5])
    tab = o.point_collection_table(g, px, py, o.PRIMAL, "m4prime")
    d = o.forcing_line(np.zeros(shp), tab, strength)
    assert abs(d.sum() * g.dx ** 2 - strength.sum()) < 1e-12
    xg, yg = g.coords(o.PRIMAL)
    assert abs((d * xg[:, None]).sum() * g.dx ** 2 - (strength * px).sum()) < 1e-12      # first moment of M4'
    # line forcing on the body's own table: flux * perimeter
    d = o.forcing_line(np.ones(shp), c.tabs[o.PRIMAL], np.full(c.N, -2.0))
    assert abs((d - 1.0).sum() * g.dx ** 2 + 2.0 * c.ds.sum()) < 1e-10
    # area forcing: dy + str * mask; whole-domain region = ones
    m = c.mask()
    s = np.random.default_rng(1).standard_normal(shp)
    dy = np.random.default_rng(2).standard_normal(shp)
    assert np.array_equal(o.forcing_area(dy, s, m), dy + s * m)
    assert np.array_equal(o.forcing_area(dy, s), o.forcing_area(dy, s, np.ones(shp)))
    assert abs(o.forcing_area(np.zeros(shp), np.full(shp, 3.0), m).sum() * g.dx ** 2 - 3.0 * np.pi) < 6e-2


# ---------------------------------------------------------------- Helmholtz decomposition (test/surface_ops.jl:319-372)
def test_helmholtz_round_trip(small_vcache):
    """The reference's own check: the masked curl / divergence of the recomposed field give back the inputs on the
    interior (1e-8 there; the restatement reaches 1e-10)."""
    c = small_vcache
    g = c.grid
    rng = np.random.default_rng(7)
    w = rng.standard_normal(o.field_shape(o.DUAL, g.NX, g.NY))
    d = rng.standard_normal(o.field_shape(o.PRIMAL, g.NX, g.NY))
    dvu, dvv = rng.standard_normal(c.N), rng.standard_normal(c.N)
    u, v = o.vecfield_helmholtz(c, w, d, dvu, dvv, (0.0, 0.0))
    w2 = c.curl_e2n(u, v)
    mw = o.helmholtz_jump(c, "cross", -1, dvu, dvv, w2)
    assert np.abs(mw[1:-1, 1:-1] - w[1:-1, 1:-1]).max() < 1e-8
    d2 = c.divergence(u, v)
    md = o.helmholtz_jump(c, "dot", -1, dvu, dvv, d2)
    assert np.abs(md[1:-1, 1:-1] - d[1:-1, 1:-1]).max() < 1e-8
    # the +1 / -1 conversions are inverse to each other up to rounding
    back = o.helmholtz_jump(c, "cross", +1, dvu, dvv, mw)
    assert np.abs(back - w2).max() < 1e-12 * np.abs(w2).max()
    # uniform field: grad(Vx x + Vy y) = (Vx, Vy) on the interior edges
    xg, yg = g.coords(o.PRIMAL)
    gu, gv = c.grad(1.5 * xg[:, None] - 0.5 * yg[None, :] + 0 * d)
    assert np.abs(gu[1:-1, :] - 1.5).max() < 1e-12 and np.abs(gv[:, 1:-1] + 0.5).max() < 1e-12


def test_unbounded_heat_step_conserves_point_heating():
    """heat_unbounded_step (test/literate/heatconduction-unbounded.jl recursion): with a constant point source of
    strength q the integral of T grows by q dt per step (the integrating factor and the M3 kernel both sum to one)."""
    from ilm_b200 import timemarching as tm
    g = o.Grid(64, 64, 4.0 / 62, (32, 32))
    kappa, dt = 0.005, 0.02
    tab = tm.LISKA_IFHERK
    stage_a, prev = [], 0.0
    for c in tab["c"]:
        stage_a.append(kappa / g.dx ** 2 * (c - prev) * dt)
        prev = c
    tables = {a: lgfmod.intfact_table(a, g.NX) for a in set(stage_a)}
    tabp = o.point_collection_table(g, [0.3], [-0.2], o.PRIMAL, "m3")
    T = np.zeros(o.field_shape(o.PRIMAL, g.NX, g.NY))
    for n in range(4):
        T = o.heat_unbounded_step(g, T, n * dt, dt, kappa, tab["a"], tab["c"], tables,
                                  lambda TT, t: o.forcing_line(np.zeros_like(TT), tabp, np.full(1, 5.0)))
    assert abs(T.sum() * g.dx ** 2 - 5.0 * 4 * dt) < 1e-12


# ---------------------------------------------------------------- FFT-free table forms (full-size checks of the GPU tests)
def test_table_form_schur_matches_probing():
    """oracle/direct_conv.c:ilm_oracle_table_schur (long double, no FFT) against the column-by-column
    restatement of create_RTLinvR (src/matrix_operators.jl:9-30), incl. a clipped window and a column range."""
    g = o.Grid(64, 48, 4.0 / 62, (32, 24))
    x, y, nx, ny, ds = bodies.circle(1.0, 1.4 * g.dx)
    x = x + 0.93                                             # windows clipped at the +x boundary
    G = lgfmod.lgf_table(64)
    c = o.ScalarCache(g, x, y, nx, ny, ds, G)
    S = c.create_RTLinvR()
    St = c.create_RTLinvR_table()
    assert np.abs(S - St).max() < 1e-13 * np.abs(S).max()
    blk = c.create_RTLinvR_table(scale=2.0, cols=(5, 9))
    assert np.abs(blk - 2.0 * S[:, 5:9]).max() < 1e-13 * np.abs(S).max()
    # integrating-factor table, truncated: entries beyond the table count as zero
    E = lgfmod.intfact_table(0.5, 64)
    plan = o.ConvPlan(E[:64, :48])
    tp = c.tabs[o.PRIMAL]
    Sp = np.zeros((c.N, c.N))
    for col in range(c.N):
        e = np.zeros(c.N); e[col] = 1.0
        Sp[:, col] = -o.interpolate(tp, plan.apply(o.regularize(tp, e)))
    assert np.abs(o.table_schur(tp, E[:20, :20], coef=-1.0) - Sp).max() < 1e-13 * np.abs(Sp).max()
    ER = -(o.E_matrix(tp) @ o.R_matrix(tp)).toarray()
    assert np.abs(o.table_schur(tp, np.ones((1, 1)), coef=-1.0) - ER).max() < 1e-14 * np.abs(ER).max()


def test_refined_solve_converges_to_the_operator_solution():
    """refined_solve: the fixed point solves the oracle's operator equation whatever (close) matrix preconditions it."""
    import scipy.linalg
    g = o.Grid(64, 64, 4.0 / 62, (32, 32))
    body = bodies.circle(1.0, 1.4 * g.dx)
    c = o.ScalarCache(g, *body, lgfmod.lgf_table(64))
    S = c.create_CLinvCT()
    b = np.random.default_rng(5).standard_normal(c.N)
    ref = np.linalg.solve(S, b)
    Sp = S * (1.0 + 1e-6 * np.random.default_rng(6).standard_normal(S.shape))      # a perturbed preconditioner
    mv = lambda x: -c.surface_curl_n2s(c.inverse_laplacian(c.surface_curl_s2n(x)))  # noqa: E731
    x, res = o.refined_solve(mv, b, scipy.linalg.lu_factor(Sp), tol=1e-13)
    assert res < 1e-13
    assert np.abs(S @ x - b).max() < 1e-11 * np.abs(b).max()
    assert np.abs(x - ref).max() < 1e-6 * np.abs(ref).max()                          # cond(S)-limited


def test_heat_step_table_form_equals_probing():
    g = o.Grid(48, 48, 4.0 / 46, (24, 24))
    body = bodies.circle(0.8, 1.4 * g.dx)
    c = o.ScalarCache(g, *body, lgfmod.lgf_table(48))
    from ilm_b200 import timemarching as tm
    tab_a, tab_c = tm.LISKA_IFHERK["a"], tm.LISKA_IFHERK["c"]
    dt = g.dx ** 2
    stage_a, prev = [], 0.0
    for cc in tab_c:
        stage_a.append(1.0 / g.dx ** 2 * (cc - prev) * dt)
        prev = cc
    tables = {a: lgfmod.intfact_table(a, 48) for a in set(stage_a)}
    T0 = np.zeros(o.field_shape(o.PRIMAL, 48, 48))
    T1, s1 = o.heat_ifherk_step(c, T0, 0.0, dt, 1.0, tab_a, tab_c, tables, 0.0, 1.0)
    T2, s2 = o.heat_ifherk_step(c, T0, 0.0, dt, 1.0, tab_a, tab_c, tables, 0.0, 1.0, schur="table")
    assert np.abs(T1 - T2).max() < 1e-12 * np.abs(T1).max()


@pytest.mark.parametrize("clipped", [False, True])
@pytest.mark.parametrize("scaling", [o.GRID_SCALING, o.INDEX_SCALING])
def test_schur_builders_are_symmetric_up_to_the_column_weight(scaling, clipped):
    """What the symmetric Schur build of the CUDA path rests on (csrc/ilm_api.cu, schur_symm_build): with wR = e * wgt and
    wE = e (the same DDF window on both sides) create_RTLinvR (src/matrix_operators.jl:9-30) is a symmetric matrix times
    diag(wgt), wgt = ds/dx^2 (GridScaling) or 1 (IndexScaling): S[k,c] wgt_k = S[c,k] wgt_c to rounding -- also when windows
    are clipped at the grid boundary (the dropped entries are dropped on both sides).  The stencil builders (create_CLinvCT,
    create_GLinvD, create_GLinvD_cross) share the property for bodies away from the boundary only (a clipped edge window and
    the truncated stencil next to it do not commute), which is why the CUDA path mirrors create_RTLinvR alone."""
    g = ilm_b200.PhysicalGrid(60, 52, 4.0 / 58, (30, 26))
    x, y, nx, ny, ds = ilm_b200.bodies.ellipse(0.9, 0.5, 1.4 * g.dx, center=(0.1, -0.05))
    body = (x + 0.95, y, nx, ny, ds) if clipped else (x, y, nx, ny, ds)  # windows clipped at +x
    G = ilm_b200.lgf.lgf_table(64)
    oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body, G, scaling=scaling)
    w = ds / g.dx ** 2 if scaling == o.GRID_SCALING else np.ones_like(ds)
    names = ("create_RTLinvR",) if clipped else ("create_RTLinvR", "create_CLinvCT", "create_GLinvD", "create_GLinvD_cross")
    for name in names:
        S = getattr(oc, name)()
        T = S / w[None, :]
        assert np.abs(T - T.T).max() <= 1e-13 * np.abs(T).max(), name

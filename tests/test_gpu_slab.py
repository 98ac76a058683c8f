"""Slab-decomposed inverse Laplacian (csrc/ilm_slab.cu, shard.SlabLaplacian):
the three CUDA stages with their pack / unpack, run for P virtual ranks on ONE GPU (the all-to-all is
done by slicing the packed buffers exactly as all_to_all_single would), and through torch.distributed
with a single-rank NCCL group.  Compared with the single-GPU path and with the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L
from ilm_b200 import shard

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def relerr(a, b):
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


virtual_slab_solve = shard.slab_solve_virtual_ranks


@pytest.mark.parametrize("NX,NY", [(128, 128), (91, 67), (40, 600), (600, 40)])
@pytest.mark.parametrize("P", [1, 2, 3])
def test_slab_virtual_ranks_match_single_gpu(NX, NY, P):
    g = ilm.PhysicalGrid(NX, NY, 0.05, (NX // 2, NY // 2))
    G = ilm.lgf.lgf_table(max(NX, NY))
    body = ilm.bodies.circle(0.4, 0.07)
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
    oc = o.ScalarCache(o.Grid(NX, NY, 0.05, (NX // 2, NY // 2)), *body[:5], G)
    rng = np.random.default_rng(NX + NY)
    w = rng.standard_normal(g.layout_shape(L.NODES_PRIMAL))
    (got,) = virtual_slab_solve(cache, [L.NODES_PRIMAL], [w], P)
    assert relerr(got, oc.inverse_laplacian(w.copy())) < RTOL
    # an Edges pair (two layouts with different row counts ride one complex transform)
    u = rng.standard_normal(g.layout_shape(L.XEDGES))
    v = rng.standard_normal(g.layout_shape(L.YEDGES))
    gu, gv = virtual_slab_solve(cache, [L.XEDGES, L.YEDGES], [u, v], P)
    ru, rv = oc.inverse_laplacian(u.copy()), oc.inverse_laplacian(v.copy())
    assert relerr(gu, ru) < RTOL and relerr(gv, rv) < RTOL


def test_slab_full_size_4096_virtual_ranks():
    """Full BASELINE size: a 4096^2 solve split over 4 virtual ranks is bit-identical to the
    single-GPU solve (same kernels on the same values) and inverts the 5-point Laplacian."""
    import torch
    g = ilm.PhysicalGrid.centered(4096)
    cache = ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g, device=True)
    shape = g.layout_shape(L.NODES_PRIMAL)
    w = np.random.default_rng(0).standard_normal(shape)
    (got,) = virtual_slab_solve(cache, [L.NODES_PRIMAL], [w], 4)
    single = ilm.Nodes(ilm.Primal, g, device=True).set(w)
    ilm.inverse_laplacian(single, cache)
    assert np.array_equal(got, single.array())
    lap = ilm.Nodes(ilm.Primal, g, device=True)
    ilm.laplacian(lap, ilm.Nodes(ilm.Primal, g, device=True).set(got), cache)
    assert relerr(lap.array()[1:-1, 1:-1], w[1:-1, 1:-1]) < 1e-9
    del torch


def test_slab_through_torch_distributed_single_rank():
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29871 + os.getpid() % 1000))
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        g = ilm.PhysicalGrid.centered(128)
        G = ilm.lgf.lgf_table(128)
        body = ilm.bodies.circle(1.0, 1.4 * g.dx)
        cache = ilm.SurfaceScalarCache(body, g, lgf_table=G, device=True)
        oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G)
        slab = shard.SlabLaplacian(cache, L.NODES_PRIMAL)
        w = np.random.default_rng(9).standard_normal(g.layout_shape(L.NODES_PRIMAL))
        mine = slab.scatter(w)
        slab.inverse_laplacian(mine)
        torch.cuda.synchronize()
        assert relerr(slab.gather(mine), oc.inverse_laplacian(w.copy())) < RTOL
    finally:
        if created:
            dist.destroy_process_group()

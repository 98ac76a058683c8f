"""Helmholtz decomposition of a vector field on the grid: host-side mirror of src/helmholtz.jl over the C ABI.

    v = curl(psi e_z) + grad(phi) + vp,   L psi = -(masked curl v) - R n x [v],   L phi = masked div v + R n . [v]

The reference composes these from whole-field sweeps (fill!, regularize!, .+=, two inverse Laplacians, curl!, grad!,
two sums) with three extra caches of temporaries.  Here the jump terms are accumulated on the cells under the DDF
windows only, both inverse Laplacians ride ONE complex transform (ilm_helmholtz_potentials) and the recomposition is
one sweep (ilm_vecfield_from_potentials); the extra caches hold nothing and exist for signature compatibility.
Function names and argument order follow the reference (trailing `!` dropped)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import api as A
from ._lib import MethodError


class ScalarPotentialCache:
    """ScalarPotentialCache(base_cache) (src/helmholtz.jl:3-8, 32-38): the base cache must carry vector data."""

    def __init__(self, base_cache):
        A._need_vector(base_cache, "ScalarPotentialCache")
        self.base_cache = base_cache


class VectorPotentialCache:
    """VectorPotentialCache(base_cache) (:10-15, 47-53)."""

    def __init__(self, base_cache):
        self.base_cache = base_cache


class VectorFieldCache:
    """VectorFieldCache(base_cache) (:17-24, 62-70)."""

    def __init__(self, base_cache):
        self.wcache = VectorPotentialCache(base_cache)
        self.dcache = ScalarPotentialCache(base_cache)


def _p(d):
    return None if d is None else A._ptr(d.data)


def _dual(x, what):
    return A._expect_nodes(x, A.Dual, what)


def _primal(x, what):
    return A._expect_nodes(x, A.Primal, what)


def _jump(cache, dv, what):
    A._expect(dv, A.VectorData, what)
    cache._check_n(dv, 2 * cache.N, what)
    return dv


# ---- vector potential (streamfunction)
def vectorpotential_from_masked_curlv(psi, masked_curlv, dv, base_cache, wcache=None):
    """vectorpotential_from_masked_curlv!(psi, masked_curlv, dv, base_cache, wcache) (:84-96)."""
    w = "vectorpotential_from_masked_curlv"
    L.check(base_cache._lib.ilm_helmholtz_potentials(base_cache._plan, _p(_dual(masked_curlv, w)), None, _p(_jump(base_cache, dv, w)),
                                                     _p(_dual(psi, w)), None))
    return psi


def vectorpotential_from_curlv(psi, curlv, base_cache):
    """vectorpotential_from_curlv!(psi, curlv, base_cache) (:114-124): psi = -L^-1 curlv."""
    w = "vectorpotential_from_curlv"
    L.check(base_cache._lib.ilm_helmholtz_potentials(base_cache._plan, _p(_dual(curlv, w)), None, None, _p(_dual(psi, w)), None))
    return psi


def vecfield_from_vectorpotential(v, psi, base_cache):
    """vecfield_from_vectorpotential!(v, psi, base_cache) (:130-134): v = curl(psi e_z)."""
    A._expect(v, A.Edges, "vecfield_from_vectorpotential")
    L.check(base_cache._lib.ilm_vecfield_from_potentials(base_cache._plan, _p(_dual(psi, "vecfield_from_vectorpotential")), None, None, _p(v)))
    return v


def _jump_add(op, sign, out, field, dv, cache, what, kind):
    kind(out, what), kind(field, what)
    L.check(cache._lib.ilm_helmholtz_jump_add(cache._plan, op, sign, _p(_jump(cache, dv, what)), _p(field), _p(out)))
    return out


def masked_curlv_from_curlv_masked(masked_curlv, curlv, dv, base_cache, wcache=None):
    """masked_curlv_from_curlv_masked! (:149-164): masked curl = curl of the masked field - R n x [v]."""
    return _jump_add(0, -1, masked_curlv, curlv, dv, base_cache, "masked_curlv_from_curlv_masked", _dual)


def curlv_masked_from_masked_curlv(curlv, masked_curlv, dv, base_cache, wcache=None):
    """curlv_masked_from_masked_curlv! (:171-185)."""
    return _jump_add(0, +1, curlv, masked_curlv, dv, base_cache, "curlv_masked_from_masked_curlv", _dual)


# ---- scalar potential
def scalarpotential_from_masked_divv(phi, masked_divv, dv, base_cache, dcache=None):
    """scalarpotential_from_masked_divv!(phi, masked_divv, dv, base_cache, dcache) (:186-201)."""
    w = "scalarpotential_from_masked_divv"
    A._need_vector(base_cache, w)
    L.check(base_cache._lib.ilm_helmholtz_potentials(base_cache._plan, None, _p(_primal(masked_divv, w)), _p(_jump(base_cache, dv, w)),
                                                     None, _p(_primal(phi, w))))
    return phi


def scalarpotential_from_divv(phi, divv, base_cache):
    """scalarpotential_from_divv!(phi, divv, base_cache) (:216-219): phi = L^-1 divv."""
    w = "scalarpotential_from_divv"
    L.check(base_cache._lib.ilm_helmholtz_potentials(base_cache._plan, None, _p(_primal(divv, w)), None, None, _p(_primal(phi, w))))
    return phi


def vecfield_from_scalarpotential(v, phi, base_cache):
    """vecfield_from_scalarpotential!(v, phi, base_cache) (:228-232): v = grad(phi)."""
    A._expect(v, A.Edges, "vecfield_from_scalarpotential")
    L.check(base_cache._lib.ilm_vecfield_from_potentials(base_cache._plan, None, _p(_primal(phi, "vecfield_from_scalarpotential")), None, _p(v)))
    return v


def masked_divv_from_divv_masked(masked_divv, divv, dv, base_cache, dcache=None):
    """masked_divv_from_divv_masked! (:240-256)."""
    return _jump_add(1, -1, masked_divv, divv, dv, base_cache, "masked_divv_from_divv_masked", _primal)


def divv_masked_from_masked_divv(divv, masked_divv, dv, base_cache, dcache=None):
    """divv_masked_from_masked_divv! (:263-278)."""
    return _jump_add(1, +1, divv, masked_divv, dv, base_cache, "divv_masked_from_masked_divv", _primal)


# ---- both potentials in one transform, and the full decomposition
def potentials_from_masked_fields(psi, phi, masked_curlv, masked_divv, dv, base_cache):
    """psi and phi of vecfield_helmholtz! (:293-298) in one call (one complex transform for the two solves)."""
    w = "potentials_from_masked_fields"
    A._need_vector(base_cache, w)
    L.check(base_cache._lib.ilm_helmholtz_potentials(base_cache._plan, _p(_dual(masked_curlv, w)), _p(_primal(masked_divv, w)),
                                                     _p(_jump(base_cache, dv, w)), _p(_dual(psi, w)), _p(_primal(phi, w))))
    return psi, phi


def vecfield_uniformvecfield(v, Vx, Vy, base_cache):
    """vecfield_uniformvecfield!(v, Vx, Vy, base_cache) (:344-348)."""
    A._expect(v, A.Edges, "vecfield_uniformvecfield")
    v.data[: v.nu] = float(Vx)              # numpy array or CUDA tensor: same slice assignment
    v.data[v.nu:] = float(Vy)
    return v


def vecfield_helmholtz(v, masked_curlv, masked_divv, dv, vp, base_cache, veccache=None):
    """vecfield_helmholtz!(v, curlv, divv, dv, vp, base_cache, veccache) (:285-310); vp: Edges, None or a tuple
    (Vx, Vy) for a uniform field."""
    w = "vecfield_helmholtz"
    A._need_vector(base_cache, w)
    A._expect(v, A.Edges, w)
    if isinstance(vp, tuple):
        vp = vecfield_uniformvecfield(base_cache.zeros_grid(), vp[0], vp[1], base_cache)
    elif vp is not None:
        A._expect(vp, A.Edges, w)
    L.check(base_cache._lib.ilm_vecfield_helmholtz(base_cache._plan, _p(_dual(masked_curlv, w)), _p(_primal(masked_divv, w)),
                                                   _p(_jump(base_cache, dv, w)), _p(vp), _p(v)))
    return v


def _coordinate_field(field, g, axis):
    x, y = g.coordinates(field.layout)
    vals = np.broadcast_to(x[:, None] if axis == 0 else y[None, :], field.shape)
    return np.asarray(vals, dtype=np.float64)


def vectorpotential_uniformvecfield(psi, Vx, Vy, base_cache):
    """vectorpotential_uniformvecfield! (:317-324): psi = Vx y - Vy x on the dual nodes."""
    _dual(psi, "vectorpotential_uniformvecfield")
    g = base_cache.g
    t = Vx * _coordinate_field(psi, g, 1)
    return psi.set(t - Vy * _coordinate_field(psi, g, 0))


def scalarpotential_uniformvecfield(phi, Vx, Vy, base_cache):
    """scalarpotential_uniformvecfield! (:331-338): phi = Vx x + Vy y on the primal nodes."""
    _primal(phi, "scalarpotential_uniformvecfield")
    g = base_cache.g
    t = Vx * _coordinate_field(phi, g, 0)
    return phi.set(t + Vy * _coordinate_field(phi, g, 1))

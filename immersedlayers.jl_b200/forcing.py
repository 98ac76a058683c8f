"""Forcing regions (areas, lines, points): host-side mirror of src/forcing.jl over the C ABI.

The reference bundles a shape, a transform and a model function into an
`AreaForcingModel` / `LineForcingModel` / `PointForcingModel` (src/forcing.jl:14-138), builds a
region cache for each on the base cache's grid with the base cache's Laplacian
(`AreaRegionCache`, `LineRegionCache`, `PointRegionCache`, :144-257) and applies them with
`apply_forcing!` (:375-515).  Here the region caches are child plans of the base cache's plan
(ilm_plan_create_shared: same multipliers and scratch, own points and DDF tables) and the three
`_apply_forcing!` bodies are one library call each:

    area   dy .+= str .* mask            ilm_forcing_area_add   (one fused sweep)
    line   dy .+= R str                  ilm_forcing_line_add   (only the cells under the DDF windows)
    point  dy .+= regop(str)             ilm_forcing_line_add on a weight-1 point collection

Model functions keep the reference's in-place signature `fcn(str, state, t, region_cache, phys_params)`;
`str` is a Nodes / Edges (area) or ScalarData / VectorData (line, point) container whose `.data` is a
numpy array (host mode) or a CUDA tensor (device-resident mode).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import api as A
from ._lib import DimensionMismatch, MethodError


# --------------------------------------------------------------------------
# rigid transforms of a body tuple (RigidBodyTools.MotionTransform{2} as data)
# --------------------------------------------------------------------------
class RigidTransform:
    """`RigidTransform((x, y), angle)` (test/surface_ops.jl:459): x' = R(angle) x + (x, y).  `tr(body)` moves a
    body tuple (points and normals), `tr(xs, ys)` moves point coordinates; `a * b` composes (b first) and
    `inv()` inverts, which is all `_update_shape!` needs (src/forcing.jl:314-325)."""

    def __init__(self, translation=(0.0, 0.0), angle=0.0):
        self.x, self.y = float(translation[0]), float(translation[1])
        self.angle = float(angle)

    def _rot(self, u, v):
        c, s = np.cos(self.angle), np.sin(self.angle)
        return c * u - s * v, s * u + c * v

    def __call__(self, *args):
        if len(args) == 2:
            u, v = self._rot(np.asarray(args[0], float), np.asarray(args[1], float))
            return u + self.x, v + self.y
        (body,) = args
        x, y = self(body[0], body[1])
        nx, ny = self._rot(np.asarray(body[2], float), np.asarray(body[3], float))
        return (x, y, nx, ny, np.asarray(body[4], float).copy()) + tuple(body[5:])

    def __mul__(self, other):
        tx, ty = self(np.array([other.x]), np.array([other.y]))
        return RigidTransform((tx[0], ty[0]), self.angle + other.angle)

    def inv(self):
        t = RigidTransform((0.0, 0.0), -self.angle)
        tx, ty = t._rot(np.array([-self.x]), np.array([-self.y]))
        return RigidTransform((tx[0], ty[0]), -self.angle)


MotionTransform = RigidTransform


# --------------------------------------------------------------------------
# spatial fields (CartesianGrids.SpatialGaussian / GeneratedField as data)
# --------------------------------------------------------------------------
class SpatialGaussian:
    """`SpatialGaussian(sx, sy, x0, y0, A)` (test/surface_ops.jl:534): A/(pi sx sy) exp(-((x-x0)/sx)^2 - ((y-y0)/sy)^2),
    each factor cut off at 6 radii.  [UPSTREAM-RECALL: CartesianGrids' Gaussian(sigma, x0, A) = A/(sqrt(pi) sigma)
    exp(-r^2); no reference test pins the values -- the field is plan-time input data.]"""

    def __init__(self, sx, sy, x0, y0, A):
        self.sx, self.sy, self.x0, self.y0, self.A = float(sx), float(sy), float(x0), float(y0), float(A)

    def __call__(self, x, y):
        rx = (np.asarray(x, float) - self.x0) / self.sx
        ry = (np.asarray(y, float) - self.y0) / self.sy
        gx = np.where(np.abs(rx) < 6.0, np.exp(-rx * rx), 0.0) * (self.A / np.sqrt(np.pi) / self.sx)
        gy = np.where(np.abs(ry) < 6.0, np.exp(-ry * ry), 0.0) * (1.0 / np.sqrt(np.pi) / self.sy)
        return gx * gy


class GeneratedField:
    """`GeneratedField(field_prototype, s, g)` (src/forcing.jl:293-295): the spatial field(s) evaluated once on the
    coordinates of the prototype's layout; calling it returns the grid data.  One field for Nodes, a pair for Edges."""

    def __init__(self, prototype, fields, g):
        if isinstance(prototype, A.Edges):
            fu, fv = fields if isinstance(fields, (list, tuple)) else (fields, fields)
            xu, yu = g.coordinates(L.XEDGES)
            xv, yv = g.coordinates(L.YEDGES)
            vals = np.concatenate([np.asarray(fu(xu[:, None], yu[None, :])).ravel(order="F"),
                                   np.asarray(fv(xv[:, None], yv[None, :])).ravel(order="F")])
            self.field = A.Edges(g, device=A._is_torch(prototype.data))
        else:
            x, y = g.coordinates(prototype.layout)
            vals = np.asarray(fields(x[:, None], y[None, :])).ravel(order="F")
            self.field = A.Nodes(prototype.celltype, g, device=A._is_torch(prototype.data))
        self.field.set(vals)

    def __call__(self):
        return self.field


# --------------------------------------------------------------------------
# forcing models (src/forcing.jl:14-138)
# --------------------------------------------------------------------------
class _ForcingModel:
    def __init__(self, *args, **kwargs):
        if len(args) == 1:                     # AreaForcingModel(fcn): the whole domain (:74)
            shape, transform, fcn = None, None, args[0]
        elif len(args) == 2:                   # (shape, fcn): identity transform (:28-29)
            shape, transform, fcn = args[0], RigidTransform(), args[1]
        elif len(args) == 3:
            shape, transform, fcn = args
        else:
            raise MethodError(f"{type(self).__name__}(shape, [transform,] model_function)")
        if not callable(fcn):
            raise MethodError("the model function must be callable")
        self.shape, self.transform, self.fcn, self.kwargs = shape, transform, fcn, dict(kwargs)


class AreaForcingModel(_ForcingModel):
    """AreaForcingModel(shape, transform, model_function!; ddftype, spatialfield) / AreaForcingModel(model_function!)."""


class LineForcingModel(_ForcingModel):
    """LineForcingModel(shape, transform, model_function!; ddftype)."""


class PointForcingModel(_ForcingModel):
    """PointForcingModel(pts, [transform,] model_function!; ddftype): pts = (x, y) arrays, or a
    point_function(state, t, region_cache, phys_params) returning them (:97-138)."""


# --------------------------------------------------------------------------
# region caches (src/forcing.jl:144-257)
# --------------------------------------------------------------------------
def _child_cache(shape, cache, **kwargs):
    cls = A.SurfaceVectorCache if A._is_vector(cache) else A.SurfaceScalarCache
    kw = dict(scaling=cache.scaling, ddftype=cache.ddftype, device=cache.device, parent=cache)
    kw.update({k: v for k, v in kwargs.items() if k in ("scaling", "ddftype")})
    return cls(shape, cache.g, **kw)


def _empty_body():
    z = np.zeros(0)
    return z, z, z, z, z


class AreaRegionCache:
    """AreaRegionCache(shape, cache) / AreaRegionCache(cache) (:144-157, 201-224, 251-255): `mask` of the shape on
    the base cache's grid (ones for the whole-domain region), the strength field `str`, the optional generated field."""

    def __init__(self, *args, spatialfield=None, **kwargs):
        shape, cache = (None, args[0]) if len(args) == 1 else args
        self.cache = _child_cache(_empty_body() if shape is None else shape, cache, **kwargs)
        self.whole_domain = shape is None
        self.mask = A.mask(self.cache)
        self.str = self.cache.zeros_grid()
        self.spatialfield = spatialfield
        self.generated_field = None if spatialfield is None else GeneratedField(self.str, spatialfield, cache.g)


class LineRegionCache:
    """LineRegionCache(shape, cache) (:160-165, 226-232): arc-length coordinates `s`, strength `str`."""

    def __init__(self, shape, cache, **kwargs):
        self.cache = _child_cache(shape, cache, **kwargs)
        ds = np.asarray(shape[4], float)
        first = np.asarray(shape[5]) if len(shape) > 5 else np.array([0, ds.shape[0]])
        s = np.zeros_like(ds)
        for b in range(len(first) - 1):            # midpoint arc coordinate, restarted on every body
            d = ds[first[b]:first[b + 1]]
            s[first[b]:first[b + 1]] = np.cumsum(d) - 0.5 * d
        self.s = s
        self.str = self.cache.zeros_surface()


class PointCollectionCache(A.SurfaceScalarCache):
    """ScalarPointCollectionCache / VectorPointCollectionCache (src/cache.jl:269-292): a weight-1, non-symmetric
    Regularize on the points (wR = ddf ddf / dx^2), no normals."""

    def __init__(self, pts, cache, ddftype=None, **_):
        x, y = np.asarray(pts[0], float), np.asarray(pts[1], float)
        if x.shape != y.shape:
            raise DimensionMismatch("point coordinates must have equal length")
        z = np.zeros_like(x)
        self.kind = "vector" if A._is_vector(cache) else "scalar"
        super().__init__((x, y, z, z, np.ones_like(x)), cache.g, scaling=A.GridScaling,
                         ddftype=ddftype or cache.ddftype, device=cache.device, parent=cache)

    def zeros_grid(self):
        return A.Edges(self.g, device=self.device) if self.kind == "vector" else A.Nodes(A.Primal, self.g, device=self.device)

    def zeros_surface(self):
        return (A.VectorData if self.kind == "vector" else A.ScalarData)(self.N, device=self.device)


class PointRegionCache:
    """PointRegionCache(pts, cache; ddftype) (:167-173, 234-247); pts may be a point function (empty collection)."""

    def __init__(self, pts, cache, **kwargs):
        if callable(pts):
            pts = (np.zeros(0), np.zeros(0))
        self.base = cache
        self.cache = PointCollectionCache(pts, cache, **kwargs)
        self.str = self.cache.zeros_surface()


def mask(ar):
    return ar.mask


def arcs(lr):
    return lr.s


def points(pr):
    return pr.cache.points()


# --------------------------------------------------------------------------
# model + region (src/forcing.jl:262-345)
# --------------------------------------------------------------------------
class ForcingModelAndRegionItem:
    def __init__(self, region_cache, shape, transform, fcn, kwargs):
        self.region_cache, self.shape, self.transform, self.fcn, self.kwargs = region_cache, shape, transform, fcn, kwargs


def _placed(shape, transform, Xi_to_ref):
    """_update_shape! (:314-325): the shape moved by transform * inv(Xi_to_ref)."""
    if shape is None or callable(shape) or transform is None:
        return shape
    full = transform * Xi_to_ref.inv()
    if len(shape) == 2:
        return full(shape[0], shape[1])
    return full(shape)


def _model_and_region(model, Xi_to_ref, cache):
    if isinstance(model, ForcingModelAndRegionItem):
        kind = type(model.region_cache)
    else:
        kind = {AreaForcingModel: AreaRegionCache, LineForcingModel: LineRegionCache, PointForcingModel: PointRegionCache}.get(type(model))
        if kind is None:
            raise MethodError(f"not a forcing model: {type(model).__name__}")
    shape = _placed(model.shape, model.transform, Xi_to_ref)
    if kind is AreaRegionCache and shape is None:
        rc = AreaRegionCache(cache, **model.kwargs)
    else:
        rc = kind(shape, cache, **model.kwargs)
    return ForcingModelAndRegionItem(rc, model.shape, model.transform, model.fcn, model.kwargs)


def ForcingModelAndRegion(models, cache, Xi_to_ref=None):
    """ForcingModelAndRegion(model or list of models, cache) -> list (:327-345); `None` gives an empty list."""
    Xi_to_ref = Xi_to_ref or RigidTransform()
    if models is None:
        return []
    if not isinstance(models, (list, tuple)):
        models = [models]
    return [_model_and_region(m, Xi_to_ref, cache) for m in models]


# --------------------------------------------------------------------------
# application (src/forcing.jl:375-515)
# --------------------------------------------------------------------------
def _assign_check(strdata, n):
    if len(strdata) != n:
        raise DimensionMismatch(f"forcing strength has {len(strdata)} entries, expected {n}")


def _apply_one(dy, y, t, f, phys_params, Xi_to_ref):
    rc = f.region_cache
    if isinstance(rc, AreaRegionCache):
        rc.str.fill(0.0)
        f.fcn(rc.str, y, t, rc, phys_params)
        _assign_check(rc.str, len(dy))
        m = None if rc.whole_domain and rc.cache.N == 0 else rc.mask
        L.check(rc.cache._lib.ilm_forcing_area_add(rc.cache._plan, dy.layout, A._ptr(rc.str.data),
                                                   A._ptr(m.data) if m is not None else None, A._ptr(dy.data)))
        return dy
    if isinstance(rc, PointRegionCache) and callable(f.shape):
        # instantaneous point collection from the point function (:496-515)
        pts = f.shape(y, t, rc, phys_params)
        pts = _placed((np.asarray(pts[0], float), np.asarray(pts[1], float)), f.transform, Xi_to_ref)
        new_rc = PointRegionCache(pts, rc.base, **f.kwargs)
        f.fcn(new_rc.str, y, t, rc, phys_params)
        rc = new_rc
    else:
        rc.str.fill(0.0)
        f.fcn(rc.str, y, t, rc, phys_params)
    ncomp = 2 if dy.layout == L.EDGES else 1
    _assign_check(rc.str, ncomp * rc.cache.N)
    L.check(rc.cache._lib.ilm_forcing_line_add(rc.cache._plan, dy.layout, A._ptr(rc.str.data), A._ptr(dy.data)))
    return dy


def apply_forcing(out, y, x, t, fr, phys_params=None, motions=None, base_cache=None):
    """apply_forcing!(out, y, x, t, fr, phys_params, motions, base_cache) (:375-396): `out` is zeroed, then every
    model adds its contribution.  Without motions the given region caches are used as they are (:399)."""
    if isinstance(fr, ForcingModelAndRegionItem):
        fr = [fr]
    Xi_to_ref = RigidTransform()
    if motions is not None:                       # (:401-414) regenerate about the reference body's axes
        Xi_to_ref = motions
        fr = ForcingModelAndRegion(fr, base_cache, Xi_to_ref=Xi_to_ref)
    out.fill(0.0)
    for f in fr:
        _apply_one(out, y, t, f, phys_params, Xi_to_ref)
    return out

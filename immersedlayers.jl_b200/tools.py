"""Inner products, integrals and body-list helpers: host-side mirror of src/tools.jl (and of the coordinate /
ones fields of src/cache.jl:351-775).  These are bookkeeping utilities around the hot path (diagnostics such as the
added mass `-integrate(df o nrm, sys, 2)` of test/literate/neumann.jl:176), evaluated with numpy on a host copy of
the data; nothing here launches a kernel.

Grid inner products follow CartesianGrids' trapezoid weights: along a primal direction the first and last entries
weigh 1/2, along a dual direction the ghost entries weigh 0, times the cell area dx^2 (`dot(u1,u2,g) = dot(u1,u2)
volume(g)`, src/tools.jl:101-104; pinned by `integrate(ones_grid) = 16.3216` of examples/caches.ipynb)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import api as A
from ._lib import DimensionMismatch, MethodError

_SHIFT = {L.NODES_PRIMAL: (False, False), L.NODES_DUAL: (True, True), L.XEDGES: (True, False), L.YEDGES: (False, True)}


def _weights(shape, layout):
    def w1(m, dual):
        w = np.ones(m)
        w[0] = w[-1] = 0.0 if dual else 0.5
        return w
    dx_dual, dy_dual = _SHIFT[layout]
    return w1(shape[0], dx_dual)[:, None] * w1(shape[1], dy_dual)[None, :]


def _components(u, g):
    """[(array (mx, my), layout)] of a grid-data container."""
    if isinstance(u, A.Edges):
        return [(u.u, L.XEDGES), (u.v, L.YEDGES)]
    if isinstance(u, A.EdgeGradient):
        return [(u.component(i), lay) for i, lay in enumerate((L.NODES_PRIMAL, L.NODES_DUAL, L.NODES_DUAL, L.NODES_PRIMAL))]
    if isinstance(u, (A.Nodes, A.XEdges, A.YEdges)):
        return [(u.array(), u.layout)]
    raise MethodError(f"no grid inner product for {type(u).__name__}")


def _same_grid_type(u1, u2, g):
    if type(u1) is not type(u2) or getattr(u1, "celltype", None) != getattr(u2, "celltype", None):
        raise MethodError(f"dot: {type(u1).__name__} and {type(u2).__name__} are different grid-data types")
    if len(u1) != len(u2):
        raise MethodError("dot: grid data of different sizes")
    lays = [lay for _, lay in _components(u1, g)]
    expect = sum(int(np.prod(g.layout_shape(lay))) for lay in lays)
    if expect != len(u1):
        raise DimensionMismatch("dot: the grid data do not live on this grid")            # @assert (NX,NY) == size(g)


def _is_grid_data(u):
    return isinstance(u, (A.Nodes, A.Edges, A.XEdges, A.YEdges, A.EdgeGradient))


def _cache_bl(cache):
    return (None, None, None, None, None, cache.body_first)


def _cache_args(u, w, bl, i, for_integral=False):
    """The cache methods of src/cache.jl:815-909: GridScaling weighs grid data by dx^2 and surface data by ds,
    IndexScaling uses plain sums -- except integrate, which always uses the areas.  The body index may come as the
    fourth positional argument, as in dot(u1, u2, cache, i)."""
    if not isinstance(w, A.SurfaceScalarCache):
        return w, bl, i
    cache = w
    if i is None and isinstance(bl, (int, np.integer)):
        i = int(bl)
    bl = _cache_bl(cache) if i is not None else None
    if _is_grid_data(u):
        if cache.scaling == A.GridScaling or for_integral:
            return cache.g, None, None
        return A.PhysicalGrid(cache.g.NX, cache.g.NY, 1.0, cache.g.I0), None, None     # plain dot: unit cell area
    weights = cache.areas() if (cache.scaling == A.GridScaling or for_integral) else np.ones(cache.N)
    return weights, bl, i


def dot(u1, u2, w, bl=None, i=None):
    """dot(u1, u2, g) on grid data (src/tools.jl:101-104); dot(u1, u2, ds) on ScalarData / VectorData (:153-155);
    dot(u1, u2, ds, bl, i) restricted to body i (1-based) of a body list (:163-168); `bl` is the body tuple whose sixth
    entry holds the offsets of the bodies (bodies.concat).  dot(u1, u2, cache[, i]) scales as the cache does
    (src/cache.jl:826-858)."""
    w, bl, i = _cache_args(u1, w, bl, i)
    if isinstance(w, A.PhysicalGrid):
        _same_grid_type(u1, u2, w)
        tot = 0.0
        for (a, lay), (b, _) in zip(_components(u1, w), _components(u2, w)):
            tot += float(np.sum(_weights(a.shape, lay) * a * b))
        return tot * w.dx ** 2
    ds = np.asarray(w.numpy() if isinstance(w, A._Data) else w, dtype=np.float64)
    if type(u1) is not type(u2) or not isinstance(u1, (A.ScalarData, A.VectorData)):
        raise MethodError("dot: expected two ScalarData or two VectorData")
    a, b = u1.numpy(), u2.numpy()
    n = ds.shape[0]
    if a.shape[0] != b.shape[0] or a.shape[0] not in (n, 2 * n):
        raise DimensionMismatch("dot: point data and weights of different lengths")
    sl = slice(0, n) if bl is None else body_range(bl, i)
    tot = float(np.sum(a[:n][sl] * ds[sl] * b[:n][sl]))
    if isinstance(u1, A.VectorData):
        tot += float(np.sum(a[n:][sl] * ds[sl] * b[n:][sl]))
    return tot


def norm(u, w, bl=None, i=None):
    """norm(u, g) (:110) / norm(u, ds[, bl, i]) (:175-183) / norm(u, cache[, i]) (src/cache.jl:815-842)."""
    return float(np.sqrt(dot(u, u, w, bl, i)))


def ones(u, dim=None):
    """ones(u) / ones(u, dim) (:116-133, :214-231): same type as u, filled with ones (in component dim, 1-based)."""
    o = _similar(u)
    if dim is None:
        return o.fill(1.0)
    parts = _slices(o)
    if not 1 <= dim <= len(parts):
        raise MethodError(f"ones: component {dim} of {type(u).__name__}")
    o.data[parts[dim - 1]] = 1.0
    return o


def _similar(u):
    dev = A._is_torch(u.data)
    if isinstance(u, A.Nodes):
        g = _GridOf(u)
        return A.Nodes(u.celltype, g, device=dev)
    if isinstance(u, (A.Edges, A.XEdges, A.YEdges, A.EdgeGradient)):
        return type(u)(_GridOf(u), device=dev)
    if isinstance(u, A.ScalarData):
        return A.ScalarData(len(u), device=dev)
    if isinstance(u, A.VectorData):
        return A.VectorData(u.n, device=dev)
    if isinstance(u, A.TensorData):
        return A.TensorData(u.n, device=dev)
    raise MethodError(f"ones: no method for {type(u).__name__}")


class _GridOf:
    """Just enough of a PhysicalGrid to allocate a container of the same shape as `u`."""

    def __init__(self, u):
        if isinstance(u, A.Nodes):
            mx, my = u.shape
            self.NX, self.NY = (mx + 1, my + 1) if u.celltype == A.Primal else (mx, my)
        elif isinstance(u, A.Edges):
            self.NX, self.NY = u.ushape[0], u.vshape[1]
        elif isinstance(u, A.XEdges):
            self.NX, self.NY = u.shape[0], u.shape[1] + 1
        elif isinstance(u, A.YEdges):
            self.NX, self.NY = u.shape[0] + 1, u.shape[1]
        else:                                   # EdgeGradient
            self.NX, self.NY = u.shapes[1]

    def layout_shape(self, layout):
        return A.PhysicalGrid.layout_shape(self, layout)


def _slices(o):
    if isinstance(o, A.Edges):
        return [slice(0, o.nu), slice(o.nu, o.nu + o.nv)]
    if isinstance(o, A.EdgeGradient):
        return [slice(int(o.offsets[k]), int(o.offsets[k + 1])) for k in range(4)]
    if isinstance(o, A.VectorData):
        return [slice(0, o.n), slice(o.n, 2 * o.n)]
    if isinstance(o, A.TensorData):
        return [slice(k * o.n, (k + 1) * o.n) for k in range(4)]
    return [slice(0, len(o))]


def integrate(u, w, bl=None, i=None):
    """integrate(u, g) (:142-144): a number for scalar grid data, a list per component for Edges / EdgeGradient;
    integrate(u, ds[, bl, i]) (:192-206): surface integral of ScalarData (number) or VectorData (pair);
    integrate(u, cache[, i]) (src/cache.jl:868-899): the same with the cache's grid / areas, whatever its scaling."""
    w, bl, i = _cache_args(u, w, bl, i, for_integral=True)
    if isinstance(w, A.PhysicalGrid):
        comps = _components(u, w)
        vals = [float(np.sum(_weights(a.shape, lay) * a)) * w.dx ** 2 for a, lay in comps]
        return vals[0] if len(vals) == 1 else vals
    ds = np.asarray(w.numpy() if isinstance(w, A._Data) else w, dtype=np.float64)
    n = ds.shape[0]
    a = u.numpy()
    sl = slice(0, n) if bl is None else body_range(bl, i)
    if isinstance(u, A.VectorData):
        if a.shape[0] != 2 * n:
            raise DimensionMismatch("integrate: point data and weights of different lengths")
        return [float(np.sum(a[:n][sl] * ds[sl])), float(np.sum(a[n:][sl] * ds[sl]))]
    if not isinstance(u, A.ScalarData):
        raise MethodError("integrate: expected ScalarData or VectorData")
    if a.shape[0] != n:
        raise DimensionMismatch("integrate: point data and weights of different lengths")
    return float(np.sum(a[sl] * ds[sl]))


def pointwise_dot(a, b):
    """pointwise_dot(a::VectorData, b::VectorData) -> ScalarData: a.u b.u + a.v b.v."""
    A._expect(a, A.VectorData, "pointwise_dot")
    A._expect(b, A.VectorData, "pointwise_dot")
    if a.n != b.n:
        raise DimensionMismatch("pointwise_dot: different lengths")
    return A.ScalarData(a.n).set(a.u * b.u + a.v * b.v)


# ---- body lists: a body tuple (x, y, nx, ny, ds, first) with the offsets of the bodies (bodies.concat)
def body_range(bl, i):
    """Index range of body i (1-based, as in view(v, bl, i), src/tools.jl:57); bl: a body tuple or a cache."""
    if isinstance(bl, A.SurfaceScalarCache):
        bl = _cache_bl(bl)
    first = np.asarray(bl[5]) if len(bl) > 5 else np.array([0, len(bl[0])])
    if not 1 <= i <= len(first) - 1:
        raise DimensionMismatch(f"body index {i} of a list of {len(first) - 1} bodies")
    return slice(int(first[i - 1]), int(first[i]))


def view(v, bl, i):
    """view(v, bl, i): the entries of body i; ScalarData -> 1-d array view, VectorData -> (u, v) views (host data)."""
    sl = body_range(bl, i)
    if isinstance(v, A.VectorData):
        return v.u[sl], v.v[sl]
    return (v.numpy() if isinstance(v, A._Data) else np.asarray(v))[sl]


def copyto(u, v, bl, i):
    """copyto!(u, v, bl, i) (:66-88): copy the entries of body i from v (point data of the same kind, or a plain
    vector of exactly that body's length) into u; the other entries of u are left alone."""
    sl = body_range(bl, i)
    nb = sl.stop - sl.start
    parts = _slices(u)
    if isinstance(v, A._Data):
        if type(u) is not type(v) or len(u) != len(v):
            raise MethodError("copyto!: point data of different kinds")
        src = v.numpy()
        vals = [src[p][sl] for p in parts]
    else:
        vec = np.asarray(v, dtype=np.float64)
        if vec.shape[0] != nb * len(parts):
            raise DimensionMismatch("Lengths are incompatible for copyto!")                # @assert, src/tools.jl:84
        vals = [vec[k * nb:(k + 1) * nb] for k in range(len(parts))]
    host = u.numpy().copy() if A._is_torch(u.data) else u.data
    for p, val in zip(parts, vals):
        host[p][sl] = val
    if A._is_torch(u.data):
        u.set(host)
    return u


# ---- coordinate and ones fields of a cache (src/cache.jl:351-775)
def _coord(field, g, axis):
    comps = []
    for arr, lay in _components(field, g):
        x, y = g.coordinates(lay)
        comps.append(np.broadcast_to(x[:, None] if axis == 0 else y[None, :], arr.shape).ravel(order="F"))
    return field.set(np.concatenate(comps))


def x_grid(cache):
    return _coord(cache.zeros_grid(), cache.g, 0)


def y_grid(cache):
    return _coord(cache.zeros_grid(), cache.g, 1)


def x_gridcurl(cache):
    return _coord(cache.zeros_gridcurl(), cache.g, 0)


def y_gridcurl(cache):
    return _coord(cache.zeros_gridcurl(), cache.g, 1)


def x_gridgrad(cache):
    return _coord(cache.zeros_gridgrad(), cache.g, 0)


def y_gridgrad(cache):
    return _coord(cache.zeros_gridgrad(), cache.g, 1)


def ones_grid(cache):
    return cache.zeros_grid().fill(1.0)


def ones_gridcurl(cache):
    return cache.zeros_gridcurl().fill(1.0)


def ones_gridgrad(cache):
    return cache.zeros_gridgrad().fill(1.0)


def ones_surface(cache):
    return cache.zeros_surface().fill(1.0)

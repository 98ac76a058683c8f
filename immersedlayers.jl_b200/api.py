"""Host-side mirror of the ImmersedLayers.jl operator API on `BasicILMCache`
for the hot path, over the C ABI of libilm_b200.so.

Names follow the reference (the trailing `!` is dropped, the output comes
first as in Julia): `regularize(s, f, cache)` is `regularize!(s,f,cache)`
(src/surface_operators.jl:13), etc.  Data containers are thin typed wrappers
around a flat fp64 buffer in the reference's memory order (x fastest) that is
either a numpy array (host mode: the library stages H2D/D2H inside the call)
or a CUDA torch tensor (device-resident mode: zero copies).

Error behaviour mirrors the reference: a call with the wrong container type
raises `MethodError` (Julia MethodError, test/tools.jl:34), a size mismatch
raises `DimensionMismatch` (AssertionError / DimensionMismatch, test/tools.jl:35).
"""
from __future__ import annotations

import ctypes as C
import sys
import weakref

import numpy as np

from . import _lib as L
from . import lgf as _lgf
from ._lib import DimensionMismatch, IlmError, MethodError  # noqa: F401

Primal, Dual = "primal", "dual"
GridScaling, IndexScaling = "grid", "index"


# --------------------------------------------------------------------------
# grid
# --------------------------------------------------------------------------
class PhysicalGrid:
    """size(g) = (NX, NY) dual cells incl. ghosts, cellsize dx, origin I0
    (1-based index of the primal node at the physical origin).  Dimensions are
    explicit: upstream's PhysicalGrid auto-tunes them."""

    def __init__(self, NX, NY, dx, I0):
        self.NX, self.NY, self.dx = int(NX), int(NY), float(dx)
        self.I0 = (int(I0[0]), int(I0[1]))

    @classmethod
    def centered(cls, N, half_width=2.0):
        """Square domain (-h, h)^2 with N x N dual cells incl. ghosts
        (SURVEY.md section 8d synthetic grids)."""
        return cls(N, N, 2.0 * half_width / (N - 2), (N // 2, N // 2))

    def size(self):
        return self.NX, self.NY

    def cellsize(self):
        return self.dx

    def layout_shape(self, layout):
        NX, NY = self.NX, self.NY
        return {L.NODES_PRIMAL: (NX - 1, NY - 1), L.NODES_DUAL: (NX, NY),
                L.XEDGES: (NX, NY - 1), L.YEDGES: (NX - 1, NY)}[layout]

    def coordinates(self, layout):
        mx, my = self.layout_shape(layout)
        sx, sy = {L.NODES_PRIMAL: (0.0, 0.0), L.NODES_DUAL: (0.5, 0.5),
                  L.XEDGES: (0.5, 0.0), L.YEDGES: (0.0, 0.5)}[layout]
        x = (np.arange(1, mx + 1) - sx - self.I0[0]) * self.dx
        y = (np.arange(1, my + 1) - sy - self.I0[1]) * self.dx
        return x, y


# --------------------------------------------------------------------------
# data containers
# --------------------------------------------------------------------------
def _is_torch(buf):
    return type(buf).__module__.startswith("torch")


def _ptr(buf):
    if _is_torch(buf):
        return C.c_void_p(buf.data_ptr())
    return C.c_void_p(buf.ctypes.data)


_PIN_MIN = 1 << 16          # host buffers of at least this many doubles are page-locked


_PIN_POOL = {}              # size (doubles) -> released page-locked buffers, reused instead of re-pinned
_PIN_IDS = set()            # ids of the page-locked numpy buffers handed out by _host_zeros
_PIN_POOL_MAX = 1 << 28     # at most this many doubles (2 GiB) parked in the pool
_pin_pool_size = 0


def _host_zeros(n, zero=True):
    """Host buffer for host mode (zero=False: contents undefined, for outputs that are overwritten whole).  Large buffers are page-locked (through torch) so that the
    library's H2D/D2H staging runs at DMA speed instead of through the driver's bounce buffer.
    Page-locking 134 MB costs ~30 ms, so released buffers are parked in a pool (see _recycle)
    and handed out again after a memset."""
    global _pin_pool_size
    n = int(n)
    if n >= _PIN_MIN:
        free = _PIN_POOL.get(n)
        if free:
            buf = free.pop()
            _pin_pool_size -= n
            if zero:
                buf.fill(0.0)
            return buf
        try:
            import torch
            if torch.cuda.is_available():
                buf = (torch.zeros if zero else torch.empty)(n, dtype=torch.float64, pin_memory=True).numpy()
                _PIN_IDS.add(id(buf))
                weakref.finalize(buf, _PIN_IDS.discard, id(buf))     # ids are reused after a free
                return buf
        except Exception:
            pass
    return np.zeros(n, dtype=np.float64)


def _recycle(buf, extra_refs=0):
    """Park a page-locked buffer for reuse once nothing but the caller refers to it (numpy views keep
    a reference to their base, so an outstanding view blocks the reuse)."""
    global _pin_pool_size
    if not isinstance(buf, np.ndarray) or id(buf) not in _PIN_IDS:
        return
    if sys.getrefcount(buf) > 2 + extra_refs or _pin_pool_size + buf.size > _PIN_POOL_MAX:
        _PIN_IDS.discard(id(buf))
        return
    _PIN_POOL.setdefault(buf.size, []).append(buf)
    _pin_pool_size += buf.size


def _alloc(n, device, zero=True):
    if device:
        import torch
        return (torch.zeros if zero else torch.empty)(int(n), dtype=torch.float64, device="cuda")
    return _host_zeros(n, zero)


class _Data:
    """Flat fp64 buffer + layout tag."""
    layout = None

    def __init__(self, data):
        self.data = data

    def __del__(self):
        try:
            _recycle(self.data, extra_refs=1)          # self.data + the argument
        except Exception:
            pass

    def __len__(self):
        return int(self.data.shape[0]) if self.data.ndim == 1 else int(np.prod(self.data.shape))

    def numpy(self):
        if _is_torch(self.data):
            return self.data.detach().cpu().numpy()
        return self.data

    @property
    def is_complex(self):
        return self.data.is_complex() if _is_torch(self.data) else np.iscomplexobj(self.data)

    @property
    def eltype(self):
        return complex if self.is_complex else float

    def fill(self, value):
        if _is_torch(self.data):
            self.data.fill_(value)
        else:
            self.data[...] = value
        return self

    def set(self, arr):
        arr = np.asarray(arr, dtype=np.complex128 if self.is_complex else np.float64).ravel(order="F")
        if arr.shape[0] != len(self):
            raise DimensionMismatch(f"expected {len(self)} values, got {arr.shape[0]}")
        if _is_torch(self.data):
            import torch
            self.data.copy_(torch.from_numpy(np.ascontiguousarray(arr)))
        else:
            self.data[...] = arr
        return self


class Nodes(_Data):
    """Nodes{Primal} ((NX-1) x (NY-1)) or Nodes{Dual} (NX x NY)."""

    def __init__(self, celltype, grid, data=None, device=False):
        self.celltype = celltype
        self.layout = L.NODES_PRIMAL if celltype == Primal else L.NODES_DUAL
        self.shape = grid.layout_shape(self.layout)
        super().__init__(_alloc(self.shape[0] * self.shape[1], device) if data is None else data)

    def array(self):
        """(mx, my) Fortran-ordered view/copy for inspection."""
        return self.numpy().reshape(self.shape, order="F")


class Edges(_Data):
    """Edges{Primal}: data = [vec(u); vec(v)], u NX x (NY-1), v (NX-1) x NY."""
    layout = L.EDGES

    def __init__(self, grid, data=None, device=False):
        self.ushape = grid.layout_shape(L.XEDGES)
        self.vshape = grid.layout_shape(L.YEDGES)
        self.nu = self.ushape[0] * self.ushape[1]
        self.nv = self.vshape[0] * self.vshape[1]
        super().__init__(_alloc(self.nu + self.nv, device) if data is None else data)

    @property
    def u(self):
        return self.numpy()[: self.nu].reshape(self.ushape, order="F")

    @property
    def v(self):
        return self.numpy()[self.nu:].reshape(self.vshape, order="F")


class _EdgeComponent(_Data):
    def __init__(self, grid, data=None, device=False):
        self.shape = grid.layout_shape(self.layout)
        super().__init__(_alloc(self.shape[0] * self.shape[1], device) if data is None else data)

    def array(self):
        return self.numpy().reshape(self.shape, order="F")


class XEdges(_EdgeComponent):
    """XEdges{Primal}: one u component, NX x (NY-1)."""
    layout = L.XEDGES


class YEdges(_EdgeComponent):
    """YEdges{Primal}: one v component, (NX-1) x NY."""
    layout = L.YEDGES


class ScalarData(_Data):
    def __init__(self, n, data=None, device=False):
        super().__init__(_alloc(n, device) if data is None else data)


class VectorData(_Data):
    """data = [u; v]."""

    def __init__(self, n, data=None, device=False):
        self.n = int(n)
        super().__init__(_alloc(2 * self.n, device) if data is None else data)

    @property
    def u(self):
        return self.numpy()[: self.n]

    @property
    def v(self):
        return self.numpy()[self.n:]


def _expect(obj, cls, what):
    if not isinstance(obj, cls):
        names = "/".join(c.__name__ for c in (cls if isinstance(cls, tuple) else (cls,)))
        raise MethodError(f"{what}: expected {names}, got {type(obj).__name__}")
    return obj


def _expect_nodes(obj, celltype, what):
    _expect(obj, Nodes, what)
    if obj.celltype != celltype:
        raise MethodError(f"{what}: expected Nodes{{{celltype}}}, got Nodes{{{obj.celltype}}}")
    return obj


# --------------------------------------------------------------------------
# the cache (plan)
# --------------------------------------------------------------------------
class SurfaceScalarCache:
    """`SurfaceScalarCache(body, g; scaling, ddftype)` (src/cache.jl:164-182,
    205-224): body = (x, y, nx, ny, ds) arrays of the surface points, unit
    normals and arc weights (`points`, `normals`, `areas`)."""

    def __init__(self, body, g, scaling=GridScaling, ddftype="yang3", lgf_table=None, c0=None,
                 device=False, stream=None, parent=None, dtype=float):
        """parent = another cache on the same grid: the Laplacian is shared (`L = L`,
        src/forcing.jl:201-248) through ilm_plan_create_shared; the parent must outlive this cache."""
        self.g = g
        x, y, nx, ny, ds = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in body[:5]]
        N = x.shape[0]
        for a in (y, nx, ny, ds):
            if a.shape[0] != N:
                raise DimensionMismatch("points, normals and areas must have equal length")
        self.pts = (x, y)
        self.nrm = (nx, ny)
        self.areas_ = ds
        self.N = N
        # offsets of the bodies of a body list (bodies.concat), a single body otherwise (cache.bl of the reference)
        self.body_first = np.asarray(body[5], dtype=np.int64) if len(body) > 5 else np.array([0, N], dtype=np.int64)
        self.scaling = scaling
        self.ddftype = ddftype
        self.device = bool(device)
        # dtype = ComplexF64 (src/cache.jl:167,187; test/surface_ops.jl:109-119): the similar_* containers are complex
        # and every (real-linear) operator acts on the real and imaginary parts, two library calls per operator
        if dtype not in (float, complex, np.float64, np.complex128):
            raise MethodError(f"dtype must be float or complex, got {dtype!r}")
        self.dtype = complex if dtype in (complex, np.complex128) else float
        if ddftype not in L.DDF:
            raise MethodError(f"unknown ddftype {ddftype!r}")
        if parent is not None:
            if (parent.g.NX, parent.g.NY, parent.g.dx, parent.g.I0) != (g.NX, g.NY, g.dx, g.I0):
                raise DimensionMismatch("a cache that shares the Laplacian must live on the parent's grid")
            self.parent = parent                  # keeps the parent plan alive
            self.lap_factor, self.c0 = parent.lap_factor, parent.c0
            lib = L.load()
            plan = C.c_void_p()
            L.check(lib.ilm_plan_create_shared(parent._plan, N, _ptr(x), _ptr(y), _ptr(nx), _ptr(ny), _ptr(ds), L.DDF[ddftype],
                                               L.GRID_SCALING if scaling == GridScaling else L.INDEX_SCALING, C.byref(plan)))
            self._plan, self._lib = plan, lib
            return
        if lgf_table is None:
            lgf_table = _lgf.lgf_table(max(g.NX, g.NY))
        lgf_table = np.asfortranarray(lgf_table, dtype=np.float64)
        if lgf_table.shape[0] != lgf_table.shape[1] or lgf_table.shape[0] < max(g.NX, g.NY):
            raise DimensionMismatch("LGF table must be square and at least max(NX,NY)")
        # _get_laplacian (src/cache.jl:321-324): factor = 1/dx^2 for GridScaling;
        # far-field constant with DX = cellsize
        self.lap_factor = 1.0 / g.dx ** 2 if scaling == GridScaling else 1.0
        self.c0 = _lgf.lgf_c0(g.dx) if c0 is None else float(c0)
        lib = L.load()
        gs = L.ilm_grid(g.NX, g.NY, g.dx, g.I0[0], g.I0[1])
        plan = C.c_void_p()
        st = lib.ilm_plan_create(C.byref(gs), N, _ptr(x), _ptr(y), _ptr(nx), _ptr(ny), _ptr(ds),
                                 L.DDF[ddftype], L.GRID_SCALING if scaling == GridScaling else L.INDEX_SCALING,
                                 _ptr(lgf_table), lgf_table.shape[0], self.c0, self.lap_factor,
                                 C.c_void_p(stream or 0), C.byref(plan))
        L.check(st)
        self._plan = plan
        self._lib = lib

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan:
            self._lib.ilm_plan_destroy(plan)
            self._plan = None

    close = __del__

    # -- convenience (src/cache.jl:351-775, src/tools.jl:16-46)
    def __len__(self):
        return self.N

    def points(self):
        return self.pts

    def normals(self):
        return self.nrm

    def areas(self):
        return self.areas_

    def cellsize(self):
        return self.g.dx

    def zeros_grid(self):
        return Nodes(Primal, self.g, device=self.device)

    def zeros_gridcurl(self):
        return Nodes(Dual, self.g, device=self.device)

    def zeros_gridgrad(self):
        return Edges(self.g, device=self.device)

    def zeros_surface(self):
        return ScalarData(self.N, device=self.device)

    def zeros_surfacevec(self):
        return VectorData(self.N, device=self.device)

    def _cplx(self, d):
        """Container of the cache's element type (complex caches: complex zeros of the same length)."""
        if self.dtype is complex and not d.is_complex:
            if _is_torch(d.data):
                import torch
                d.data = torch.zeros(len(d), dtype=torch.complex128, device=d.data.device)
            else:
                d.data = np.zeros(len(d), dtype=np.complex128)
        return d

    def eltype(self):
        return self.dtype

    def update_points(self, body):
        """update_system (src/system.jl:26-50): new body, same L."""
        x, y, nx, ny, ds = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in body[:5]]
        L.check(self._lib.ilm_plan_update_points(self._plan, x.shape[0], _ptr(x), _ptr(y), _ptr(nx), _ptr(ny), _ptr(ds)))
        self.pts, self.nrm, self.areas_, self.N = (x, y), (nx, ny), ds, x.shape[0]
        self.body_first = np.asarray(body[5], dtype=np.int64) if len(body) > 5 else np.array([0, self.N], dtype=np.int64)

    def sync(self):
        L.check(self._lib.ilm_plan_sync(self._plan))

    # -- multi-GPU: the NCCL communicator lives inside the library (ilm_comm_init); the host language only
    #    distributes the 128-byte unique id (here: torch.distributed, any backend)
    def comm_init(self, group=None):
        """Bind an NCCL communicator over the ranks of the torch.distributed `group` to this plan."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            L.check(self._lib.ilm_comm_unique_id(_ptr(uid), 128))
        box = [uid.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        L.check(self._lib.ilm_comm_init(self._plan, _ptr(uid), 128, rank, world))
        del torch
        return self

    def comm_destroy(self):
        L.check(self._lib.ilm_comm_destroy(self._plan))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        L.check(self._lib.ilm_comm_info(self._plan, C.byref(r), C.byref(n)))
        return r.value, n.value

    def launch_count(self):
        return int(self._lib.ilm_plan_launch_count(self._plan))

    def table(self, layout):
        """(idx, wR, wE) of the DDF window table, each (N, W, W) in [k][b][a] order."""
        W = C.c_int()
        L.check(self._lib.ilm_get_table(self._plan, layout, C.byref(W), None, None, None))
        W = W.value
        idx = np.zeros((self.N, W, W), dtype=np.int64)
        wR = np.zeros((self.N, W, W))
        wE = np.zeros((self.N, W, W))
        L.check(self._lib.ilm_get_table(self._plan, layout, None, _ptr(idx), _ptr(wR), _ptr(wE)))
        return idx, wR, wE

    def add_kernel(self, table, c0=0.0, factor=1.0):
        table = np.asfortranarray(table, dtype=np.float64)
        kid = C.c_int()
        L.check(self._lib.ilm_add_kernel(self._plan, _ptr(table), table.shape[0], c0, factor, C.byref(kid)))
        return kid.value

    def _check_n(self, d, n, what):
        if len(d) != n:
            raise DimensionMismatch(f"{what}: expected length {n}, got {len(d)}")


def _grid_len(cache, layout):
    return int(cache._lib.ilm_layout_size(cache._plan, layout))


# --------------------------------------------------------------------------
# operators (reference names, `!` dropped)
# --------------------------------------------------------------------------
def regularize(s, f, cache):
    """regularize!(s::Nodes{Primal}, f::ScalarData, cache)  (src/surface_operators.jl:13)
    regularize!(v::Edges, vb::VectorData, cache)           (:36)"""
    if isinstance(s, Edges):
        _expect(f, VectorData, "regularize")
        cache._check_n(f, 2 * cache.N, "regularize")
        L.check(cache._lib.ilm_regularize(cache._plan, L.EDGES, _ptr(f.data), _ptr(s.data)))
        return s
    _expect(s, Nodes, "regularize")
    _expect(f, ScalarData, "regularize")
    cache._check_n(f, cache.N, "regularize")
    cache._check_n(s, _grid_len(cache, s.layout), "regularize")
    L.check(cache._lib.ilm_regularize(cache._plan, s.layout, _ptr(f.data), _ptr(s.data)))
    return s


def interpolate(f, s, cache):
    """interpolate!(f::ScalarData, s::Nodes, cache) (:56), (vb::VectorData, v::Edges) (:78)"""
    if isinstance(s, Edges):
        _expect(f, VectorData, "interpolate")
        cache._check_n(f, 2 * cache.N, "interpolate")
        L.check(cache._lib.ilm_interpolate(cache._plan, L.EDGES, _ptr(s.data), _ptr(f.data)))
        return f
    _expect(s, Nodes, "interpolate")
    _expect(f, ScalarData, "interpolate")
    cache._check_n(f, cache.N, "interpolate")
    cache._check_n(s, _grid_len(cache, s.layout), "interpolate")
    L.check(cache._lib.ilm_interpolate(cache._plan, s.layout, _ptr(s.data), _ptr(f.data)))
    return f


def _s2e(fn_mode, q, f, cache, what):
    _expect(q, Edges, what)
    _expect(f, ScalarData, what)
    cache._check_n(f, cache.N, what)
    L.check(cache._lib.ilm_regularize_normal(cache._plan, fn_mode, _ptr(f.data), _ptr(q.data)))
    return q


def _e2s(fn_mode, vn, q, cache, what):
    _expect(q, Edges, what)
    _expect(vn, ScalarData, what)
    cache._check_n(vn, cache.N, what)
    L.check(cache._lib.ilm_normal_interpolate(cache._plan, fn_mode, _ptr(q.data), _ptr(vn.data)))
    return vn


def regularize_normal(q, f, cache):
    """regularize_normal!(q::Edges, f::ScalarData, cache) (:98)"""
    return _s2e(L.NORMAL, q, f, cache, "regularize_normal")


def regularize_normal_cross(q, f, cache):
    """regularize_normal_cross!(q::Edges, f::ScalarData, cache) (:150)"""
    return _s2e(L.CROSS, q, f, cache, "regularize_normal_cross")


def normal_interpolate(vn, q, cache):
    """normal_interpolate!(vn::ScalarData, q::Edges, cache) (:228)"""
    return _e2s(L.NORMAL, vn, q, cache, "normal_interpolate")


def normal_cross_interpolate(vn, q, cache):
    """normal_cross_interpolate!(vn::ScalarData, q::Edges, cache) (:270)"""
    return _e2s(L.CROSS, vn, q, cache, "normal_cross_interpolate")


def divergence(p, q, cache):
    """divergence!(p::Nodes{Primal}, q::Edges, cache) (src/grid_operators.jl:25)"""
    _expect_nodes(p, Primal, "divergence")
    _expect(q, Edges, "divergence")
    L.check(cache._lib.ilm_divergence(cache._plan, _ptr(q.data), _ptr(p.data)))
    return p


def grad(q, p, cache):
    """grad!(q::Edges, p::Nodes{Primal}, cache) (src/grid_operators.jl:45)"""
    _expect(q, Edges, "grad")
    _expect_nodes(p, Primal, "grad")
    L.check(cache._lib.ilm_grad(cache._plan, _ptr(p.data), _ptr(q.data)))
    return q


def curl(a, b, cache):
    """curl!(q::Edges, s::Nodes{Dual}, cache) / curl!(w::Nodes{Dual}, q::Edges, cache)
    (src/grid_operators.jl:79,97)"""
    if isinstance(a, Edges):
        _expect_nodes(b, Dual, "curl")
        L.check(cache._lib.ilm_curl_n2e(cache._plan, _ptr(b.data), _ptr(a.data)))
    else:
        _expect_nodes(a, Dual, "curl")
        _expect(b, Edges, "curl")
        L.check(cache._lib.ilm_curl_e2n(cache._plan, _ptr(b.data), _ptr(a.data)))
    return a


def laplacian(w, s, cache):
    """laplacian!(w, s, cache) (src/grid_operators.jl:123)"""
    if type(w) is not type(s) or w.layout != s.layout:
        raise MethodError("laplacian: input and output must have the same type")
    L.check(cache._lib.ilm_laplacian(cache._plan, w.layout, _ptr(s.data), _ptr(w.data)))
    return w


def inverse_laplacian(w, cache):
    """inverse_laplacian!(w, cache) in place (src/grid_operators.jl:153)"""
    _expect(w, (Nodes, Edges), "inverse_laplacian")
    cache._check_n(w, _grid_len(cache, w.layout), "inverse_laplacian")
    L.check(cache._lib.ilm_inverse_laplacian(cache._plan, w.layout, _ptr(w.data)))
    return w


def surface_divergence(theta, f, cache):
    """surface_divergence!(theta::Nodes{Primal}, f::ScalarData, cache) (src/surface_operators.jl:530)"""
    _expect_nodes(theta, Primal, "surface_divergence")
    _expect(f, ScalarData, "surface_divergence")
    cache._check_n(f, cache.N, "surface_divergence")
    L.check(cache._lib.ilm_surface_divergence(cache._plan, L.NORMAL, _ptr(f.data), _ptr(theta.data)))
    return theta


def surface_divergence_cross(theta, f, cache):
    """surface_divergence_cross! (:679; documented operator)"""
    _expect_nodes(theta, Primal, "surface_divergence_cross")
    _expect(f, ScalarData, "surface_divergence_cross")
    L.check(cache._lib.ilm_surface_divergence(cache._plan, L.CROSS, _ptr(f.data), _ptr(theta.data)))
    return theta


def surface_grad(vn, phi, cache):
    """surface_grad!(vn::ScalarData, phi::Nodes{Primal}, cache) (:604)"""
    _expect(vn, ScalarData, "surface_grad")
    _expect_nodes(phi, Primal, "surface_grad")
    L.check(cache._lib.ilm_surface_grad(cache._plan, L.NORMAL, _ptr(phi.data), _ptr(vn.data)))
    return vn


def surface_grad_cross(vn, phi, cache):
    """surface_grad_cross! (:712)"""
    _expect(vn, ScalarData, "surface_grad_cross")
    _expect_nodes(phi, Primal, "surface_grad_cross")
    L.check(cache._lib.ilm_surface_grad(cache._plan, L.CROSS, _ptr(phi.data), _ptr(vn.data)))
    return vn


def _surface_curl(mode, a, b, cache, what):
    if isinstance(a, Nodes):
        _expect_nodes(a, Dual, what)
        _expect(b, ScalarData, what)
        L.check(cache._lib.ilm_surface_curl_s2n(cache._plan, mode, _ptr(b.data), _ptr(a.data)))
    else:
        _expect(a, ScalarData, what)
        _expect_nodes(b, Dual, what)
        L.check(cache._lib.ilm_surface_curl_n2s(cache._plan, mode, _ptr(b.data), _ptr(a.data)))
    return a


def surface_curl(a, b, cache):
    """surface_curl!(w::Nodes{Dual}, f::ScalarData, cache) (:357) /
    surface_curl!(vn::ScalarData, s::Nodes{Dual}, cache) (:412)"""
    return _surface_curl(L.NORMAL, a, b, cache, "surface_curl")


def surface_curl_cross(a, b, cache):
    """surface_curl_cross! (:467, :493)"""
    return _surface_curl(L.CROSS, a, b, cache, "surface_curl_cross")


def mask(cache):
    """mask(cache) (src/surface_operators.jl:737): 1 inside, 0 outside, on Nodes{Primal}."""
    if cache.scaling != GridScaling:
        raise MethodError("mask is only defined for GridScaling caches")
    m = cache.zeros_grid()
    L.check(cache._lib.ilm_mask(cache._plan, _ptr(m.data)))
    return m


def complementary_mask(cache):
    m = mask(cache)
    if _is_torch(m.data):
        m.data.neg_().add_(1.0)
    else:
        m.data[...] = 1.0 - m.data
    return m


# --------------------------------------------------------------------------
# matrix operators (src/matrix_operators.jl) and the surface-point solve
# --------------------------------------------------------------------------
def _matrix(cache, ncols):
    if cache.device:
        import torch
        return torch.zeros(int(ncols) * cache.N, dtype=torch.float64, device="cuda")
    return _host_zeros(int(ncols) * cache.N)


def _as_matrix(buf, N, ncols):
    if _is_torch(buf):
        return buf.view(ncols, N).t()          # column-major N x ncols view
    view = buf.reshape((N, ncols), order="F")
    if id(buf) in _PIN_IDS:
        weakref.finalize(view, _recycle, buf, 1)       # the finalizer's own reference to buf
    return view


def _schur(which, cache, scale, cols):
    c0, c1 = (0, cache.N) if cols is None else cols
    buf = _matrix(cache, c1 - c0)
    L.check(cache._lib.ilm_create_schur(cache._plan, which, float(scale), int(c0), int(c1), _ptr(buf)))
    return _as_matrix(buf, cache.N, c1 - c0)


def create_RTLinvR(cache, scale=1.0, cols=None):
    """create_RTLinvR(cache; scale) (src/matrix_operators.jl:9-30).  `cols=(c0,c1)`
    builds a column block (the multi-GPU shard)."""
    return _schur(L.RTLINVR, cache, scale, cols)


def create_CLinvCT(cache, scale=1.0, cols=None):
    """create_CLinvCT (src/matrix_operators.jl:40-61)"""
    return _schur(L.CLINVCT, cache, scale, cols)


def create_GLinvD(cache, scale=1.0, cols=None):
    """create_GLinvD (src/matrix_operators.jl:135-155)"""
    return _schur(L.GLINVD, cache, scale, cols)


def create_GLinvD_cross(cache, scale=1.0, cols=None):
    """create_GLinvD_cross (src/matrix_operators.jl:195-215)"""
    return _schur(L.GLINVD_CROSS, cache, scale, cols)


def create_nRTRn(cache, scale=1.0):
    """create_nRTRn (src/matrix_operators.jl:225-244)"""
    buf = _matrix(cache, cache.N)
    L.check(cache._lib.ilm_create_nRTRn(cache._plan, float(scale), _ptr(buf)))
    return _as_matrix(buf, cache.N, cache.N)


def create_surface_filter(cache):
    """create_surface_filter (src/matrix_operators.jl:254-268)"""
    buf = _matrix(cache, cache.N)
    L.check(cache._lib.ilm_create_surface_filter(cache._plan, _ptr(buf)))
    return _as_matrix(buf, cache.N, cache.N)


def _colmajor_buffer(A):
    """Return (flat column-major buffer, n) for a square matrix given as numpy
    (any order) or as the torch column-major view produced by _as_matrix."""
    if _is_torch(A):
        n = A.shape[0]
        At = A.t()
        if not At.is_contiguous():
            At = At.contiguous()
        return At.reshape(-1), n
    Af = np.asfortranarray(A, dtype=np.float64)
    return Af.reshape(-1, order="F"), Af.shape[0]


class LU:
    """lu(S): dense LU with partial pivoting on the GPU (LAPACK getrf semantics).  The factors stay on
    the device also for a host matrix (one H2D of S; `solve` then moves only the right-hand side);
    `.lu` / `.ipiv` give them back in the memory space of the input."""

    def __init__(self, A, stream=None):
        import torch
        buf, n = _colmajor_buffer(A)
        self._host = not _is_torch(buf)
        self._lu = torch.from_numpy(np.ascontiguousarray(buf)).cuda() if self._host else buf.clone()
        self._ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
        self.n = n
        self.stream = stream
        L.check(L.load().ilm_dense_factor(n, _ptr(self._lu), _ptr(self._ipiv), C.c_void_p(stream or 0)))

    @property
    def lu(self):
        return self._lu.cpu().numpy() if self._host else self._lu

    @property
    def ipiv(self):
        return self._ipiv.cpu().numpy() if self._host else self._ipiv

    def solve(self, b):
        """S \\ b, in place on a copy; b is ScalarData or a flat buffer."""
        import torch
        data = b.data if isinstance(b, _Data) else b
        if int(np.prod(data.shape)) != self.n:
            raise DimensionMismatch(f"solve: expected length {self.n}")
        if _is_torch(data) == self._host:
            raise MethodError("solve: right-hand side and factorisation must live on the same side (host/device)")
        out = torch.from_numpy(np.array(data, dtype=np.float64, copy=True)).cuda() if self._host else data.clone()
        L.check(L.load().ilm_dense_solve(self.n, _ptr(self._lu), _ptr(self._ipiv), 1, _ptr(out), C.c_void_p(self.stream or 0)))
        if self._host:
            out = out.cpu().numpy()
        return ScalarData(self.n, data=out) if isinstance(b, _Data) else out


def matvec_pow(Cm, k, s, stream=None):
    """s <- C^k s  (`s .= C^5*s`, test/literate/dirichlet.jl:124)."""
    buf, n = _colmajor_buffer(Cm)
    data = s.data if isinstance(s, _Data) else s
    if int(np.prod(data.shape)) != n:
        raise DimensionMismatch(f"matvec_pow: expected length {n}")
    L.check(L.load().ilm_dense_matvec_pow(n, _ptr(buf), int(k), _ptr(data), C.c_void_p(stream or 0)))
    return s


# --------------------------------------------------------------------------
# the north-star algorithm: Dirichlet Poisson by block LU
# --------------------------------------------------------------------------
def dirichlet_poisson(cache, fplus, fminus=None, S=None, filter_passes=0):
    """test/literate/dirichlet.jl:71-107 on the B200 path.  fplus/fminus: surface
    values of the exterior/interior Dirichlet data.  Returns (f, s, S)."""
    N = cache.N
    fplus = np.asarray(fplus, dtype=np.float64)
    fminus = np.zeros(N) if fminus is None else np.asarray(fminus, dtype=np.float64)
    d = cache.zeros_surface().set(fplus - fminus)
    fb = 0.5 * (fplus + fminus)
    fstar = cache.zeros_grid()
    surface_divergence(fstar, d, cache)
    inverse_laplacian(fstar, cache)
    if S is None:
        S = create_RTLinvR(cache)
    s = cache.zeros_surface()
    interpolate(s, fstar, cache)
    if cache.device:
        import torch
        rhs = torch.from_numpy(fb).cuda() - s.data
        sol = LU(S).solve(rhs)
        s.data.copy_(-sol)
    else:
        rhs = fb - s.data
        sol = LU(S).solve(rhs)
        s.data[...] = -sol
    f = cache.zeros_grid()
    regularize(f, s, cache)
    inverse_laplacian(f, cache)
    if cache.device:
        f.data.add_(fstar.data)
    else:
        f.data += fstar.data
    if filter_passes:
        Cm = create_surface_filter(cache)
        matvec_pow(Cm, filter_passes, s)
    return f, s, S


def dirichlet_solve(cache, fplus, fminus=None, return_S=False, want_field=True, field_rows=None):
    """`solve(prob::DirichletPoissonProblem, sys)` of test/literate/dirichlet.jl:71-107 as ONE call of the library
    (ilm_dirichlet_poisson): boundary data in, field and multiplier out; S, its LU factors and the intermediate
    fields never leave the device, and with a communicator on the plan (cache.comm_init) the Schur columns are
    sharded over its ranks inside the library.  fplus / fminus: numpy arrays or torch CUDA tensors of length N
    (the outputs live where the cache says: device=True -> torch).  Returns (f, s) or (f, s, S); want_field=False
    (a rank of a sharded solve that only needs the multiplier) returns f = None and skips the last L^-1;
    field_rows=(r0, r1) returns the grid rows [r0, r1) of the field as a flat buffer (ilm_dirichlet_poisson_rows: every rank
    of a sharded solve keeps a slab of the result)."""
    N = cache.N
    dev = cache.device

    def surf(v):
        if v is None:
            return None
        if dev and _is_torch(v):
            return v
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        if dev:
            import torch
            return torch.from_numpy(a).cuda()
        return a

    fp, fm = surf(fplus), surf(fminus)
    for v in (fp, fm):
        if v is not None and int(np.prod(v.shape)) != N:
            raise DimensionMismatch(f"dirichlet_solve: expected {N} surface values")
    mx, my = cache.g.layout_shape(L.NODES_PRIMAL)
    s = ScalarData(N, data=_alloc(N, dev, zero=False))
    S = _matrix(cache, N) if return_S else None
    if field_rows is not None:
        r0, r1 = int(field_rows[0]), int(field_rows[1])
        if not (0 <= r0 < r1 <= my):
            raise DimensionMismatch(f"dirichlet_solve: rows [{r0}, {r1}) outside the {my} rows of Nodes{{Primal}}")
        f = _alloc((r1 - r0) * mx, dev, zero=False)
        L.check(cache._lib.ilm_dirichlet_poisson_rows(cache._plan, _ptr(fp), _ptr(fm) if fm is not None else None, _ptr(f), r0, r1,
                                                      _ptr(s.data), _ptr(S) if S is not None else None))
        return (f, s, _as_matrix(S, N, N)) if return_S else (f, s)
    f = Nodes(Primal, cache.g, data=_alloc(mx * my, dev, zero=False)) if want_field else None   # written whole by the library
    L.check(cache._lib.ilm_dirichlet_poisson(cache._plan, _ptr(fp), _ptr(fm) if fm is not None else None,
                                             _ptr(f.data) if f is not None else None,
                                             _ptr(s.data), _ptr(S) if S is not None else None))
    if return_S:
        return f, s, _as_matrix(S, N, N)
    return f, s


def create_schur_sharded(cache, which="RTLinvR", scale=1.0, kernel_id=0):
    """A Schur builder of src/matrix_operators.jl with its column loop sharded over the plan's communicator
    (ilm_create_schur_sharded): the full N x N matrix on every rank.  which: RTLinvR / CLinvCT / GLinvD / GLinvD_cross."""
    code = {"RTLinvR": L.RTLINVR, "CLinvCT": L.CLINVCT, "GLinvD": L.GLINVD, "GLinvD_cross": L.GLINVD_CROSS}[which]
    A = _matrix(cache, cache.N)
    L.check(cache._lib.ilm_create_schur_sharded(cache._plan, code, int(kernel_id), float(scale), _ptr(A)))
    return _as_matrix(A, cache.N, cache.N)


# ==========================================================================
# vector-data cache: SurfaceVectorCache (src/cache.jl:184-202) and the
# VectorData / TensorData / EdgeGradient methods of the operator API.  The
# functions below re-dispatch on the argument types the way Julia's methods do.
# ==========================================================================
EDGEGRAD = 5


class TensorData(_Data):
    """data = [dudx; dudy; dvdx; dvdy] (4N)."""

    def __init__(self, n, data=None, device=False):
        self.n = int(n)
        super().__init__(_alloc(4 * self.n, device) if data is None else data)

    def component(self, i):
        return self.numpy()[i * self.n:(i + 1) * self.n]


class EdgeGradient(_Data):
    """EdgeGradient{Primal,Dual}: [dudx; dudy; dvdx; dvdy], dudx/dvdy on primal
    nodes, dudy/dvdx on dual nodes (src/cache.jl:682-690)."""
    layout = EDGEGRAD

    def __init__(self, grid, data=None, device=False):
        ps, ds_ = grid.layout_shape(L.NODES_PRIMAL), grid.layout_shape(L.NODES_DUAL)
        self.shapes = (ps, ds_, ds_, ps)
        sizes = [s[0] * s[1] for s in self.shapes]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)])
        super().__init__(_alloc(int(self.offsets[-1]), device) if data is None else data)

    def component(self, i):
        return self.numpy()[self.offsets[i]:self.offsets[i + 1]].reshape(self.shapes[i], order="F")

    def set_components(self, comps):
        return self.set(np.concatenate([np.asarray(c).ravel(order="F") for c in comps]))


class SurfaceVectorCache(SurfaceScalarCache):
    """`SurfaceVectorCache(body, g; ...)`: surface data are VectorData, grid data
    Edges{Primal}; the plan (tables of all four layouts, Ghat) is the same object
    kind as for the scalar cache."""
    kind = "vector"

    def zeros_grid(self):
        return Edges(self.g, device=self.device)

    def zeros_gridgrad(self):
        return EdgeGradient(self.g, device=self.device)

    def zeros_griddiv(self):
        return Nodes(Primal, self.g, device=self.device)

    def zeros_surface(self):
        return VectorData(self.N, device=self.device)

    def zeros_surfacescalar(self):
        return ScalarData(self.N, device=self.device)

    def zeros_surfacetensor(self):
        return TensorData(self.N, device=self.device)


SurfaceScalarCache.kind = "scalar"


def _is_vector(cache):
    return getattr(cache, "kind", "scalar") == "vector"


def _need_vector(cache, what):
    if not _is_vector(cache):
        raise MethodError(f"{what}: only defined for a SurfaceVectorCache")


_s_regularize_normal, _s_normal_interpolate = regularize_normal, normal_interpolate
_s_regularize_normal_cross, _s_normal_cross_interpolate = regularize_normal_cross, normal_cross_interpolate
_s_divergence, _s_grad = divergence, grad
_s_surface_divergence, _s_surface_grad, _s_surface_curl = surface_divergence, surface_grad, surface_curl
_s_mask = mask
_s_create_RTLinvR, _s_create_CLinvCT, _s_create_GLinvD, _s_create_nRTRn = (
    create_RTLinvR, create_CLinvCT, create_GLinvD, create_nRTRn)


def _tensor_s2g(mode, q, f, cache, what):
    _need_vector(cache, what)
    _expect(q, EdgeGradient, what)
    _expect(f, VectorData, what)
    cache._check_n(f, 2 * cache.N, what)
    L.check(cache._lib.ilm_regularize_normal_tensor(cache._plan, mode, _ptr(f.data), _ptr(q.data)))
    return q


def _tensor_g2s(mode, vn, q, cache, what):
    _need_vector(cache, what)
    _expect(q, EdgeGradient, what)
    _expect(vn, VectorData, what)
    L.check(cache._lib.ilm_normal_interpolate_tensor(cache._plan, mode, _ptr(q.data), _ptr(vn.data)))
    return vn


def regularize_normal(q, f, cache):
    """regularize_normal!(q::Edges, f::ScalarData) (:98) / (qt::EdgeGradient, v::VectorData) (:113)"""
    if isinstance(q, EdgeGradient):
        return _tensor_s2g(0, q, f, cache, "regularize_normal")
    return _s_regularize_normal(q, f, cache)


def regularize_normal_symm(q, f, cache):
    """regularize_normal_symm!(qt::EdgeGradient, v::VectorData) (:123)"""
    return _tensor_s2g(1, q, f, cache, "regularize_normal_symm")


def normal_interpolate(vn, q, cache):
    """normal_interpolate!(vn::ScalarData, q::Edges) (:228) / (tau::VectorData, A::EdgeGradient) (:242)"""
    if isinstance(q, EdgeGradient):
        return _tensor_g2s(0, vn, q, cache, "normal_interpolate")
    return _s_normal_interpolate(vn, q, cache)


def normal_interpolate_symm(vn, q, cache):
    """normal_interpolate_symm!(tau::VectorData, A::EdgeGradient) (:252)"""
    return _tensor_g2s(1, vn, q, cache, "normal_interpolate_symm")


def regularize_normal_cross(q, f, cache):
    """regularize_normal_cross!(q::Edges, f::ScalarData) (:150) / (w::Nodes{Dual}, vs::VectorData) (:172)"""
    if isinstance(q, Nodes):
        _need_vector(cache, "regularize_normal_cross")
        _expect_nodes(q, Dual, "regularize_normal_cross")
        _expect(f, VectorData, "regularize_normal_cross")
        L.check(cache._lib.ilm_regularize_normal_vs(cache._plan, 0, _ptr(f.data), _ptr(q.data)))
        return q
    return _s_regularize_normal_cross(q, f, cache)


def regularize_normal_dot(a, b, cache):
    """regularize_normal_dot!(f::Nodes{Primal}, vs::VectorData) (:188) / (v::Edges, taus::TensorData) (:209)"""
    if isinstance(a, Edges):
        _need_vector(cache, "regularize_normal_dot")
        _expect(b, TensorData, "regularize_normal_dot")
        L.check(cache._lib.ilm_regularize_normal_dot_tensor(cache._plan, _ptr(b.data), _ptr(a.data)))
        return a
    _expect_nodes(a, Primal, "regularize_normal_dot")
    _expect(b, VectorData, "regularize_normal_dot")
    L.check(cache._lib.ilm_regularize_normal_vs(cache._plan, 1, _ptr(b.data), _ptr(a.data)))
    return a


def normal_cross_interpolate(a, b, cache):
    """normal_cross_interpolate!(wn::ScalarData, v::Edges) (:270) / (vs::VectorData, s::Nodes{Dual}) (:303)"""
    if isinstance(a, VectorData):
        _need_vector(cache, "normal_cross_interpolate")
        _expect_nodes(b, Dual, "normal_cross_interpolate")
        L.check(cache._lib.ilm_normal_interpolate_vs(cache._plan, 0, _ptr(b.data), _ptr(a.data)))
        return a
    return _s_normal_cross_interpolate(a, b, cache)


def normal_dot_interpolate(a, b, cache):
    """normal_dot_interpolate!(vs::VectorData, f::Nodes{Primal}) (:320) / (taus::TensorData, v::Edges) (:335)"""
    if isinstance(a, TensorData):
        _need_vector(cache, "normal_dot_interpolate")
        _expect(b, Edges, "normal_dot_interpolate")
        L.check(cache._lib.ilm_normal_dot_interpolate_tensor(cache._plan, _ptr(b.data), _ptr(a.data)))
        return a
    _expect(a, VectorData, "normal_dot_interpolate")
    _expect_nodes(b, Primal, "normal_dot_interpolate")
    L.check(cache._lib.ilm_normal_interpolate_vs(cache._plan, 1, _ptr(b.data), _ptr(a.data)))
    return a


def divergence(p, q, cache):
    """divergence!(p::Nodes{Primal}, q::Edges) / divergence!(p::Edges, q::EdgeGradient)"""
    if isinstance(q, EdgeGradient):
        _expect(p, Edges, "divergence")
        L.check(cache._lib.ilm_divergence_tensor(cache._plan, _ptr(q.data), _ptr(p.data)))
        return p
    return _s_divergence(p, q, cache)


def grad(q, p, cache):
    """grad!(q::Edges, p::Nodes{Primal}) (src/grid_operators.jl:45) / grad!(q::EdgeGradient, p::Edges) (:61)"""
    if isinstance(q, EdgeGradient):
        _expect(p, Edges, "grad")
        L.check(cache._lib.ilm_grad_tensor(cache._plan, _ptr(p.data), _ptr(q.data)))
        return q
    return _s_grad(q, p, cache)


def surface_divergence(theta, f, cache):
    """surface_divergence!(theta::Nodes{Primal}, f::ScalarData) (:530) / (v::Edges, dv::VectorData) (:559)"""
    if isinstance(theta, Edges):
        _need_vector(cache, "surface_divergence")
        _expect(f, VectorData, "surface_divergence")
        L.check(cache._lib.ilm_vsurface_divergence(cache._plan, 0, _ptr(f.data), _ptr(theta.data)))
        return theta
    return _s_surface_divergence(theta, f, cache)


def surface_divergence_symm(theta, f, cache):
    """surface_divergence_symm!(v::Edges, dv::VectorData) (:570)"""
    _need_vector(cache, "surface_divergence_symm")
    _expect(theta, Edges, "surface_divergence_symm")
    _expect(f, VectorData, "surface_divergence_symm")
    L.check(cache._lib.ilm_vsurface_divergence(cache._plan, 1, _ptr(f.data), _ptr(theta.data)))
    return theta


def surface_grad(vn, phi, cache):
    """surface_grad!(vn::ScalarData, phi::Nodes{Primal}) (:604) / (tau::VectorData, v::Edges) (:632)"""
    if isinstance(phi, Edges):
        _need_vector(cache, "surface_grad")
        _expect(vn, VectorData, "surface_grad")
        L.check(cache._lib.ilm_vsurface_grad(cache._plan, 0, _ptr(phi.data), _ptr(vn.data)))
        return vn
    return _s_surface_grad(vn, phi, cache)


def surface_grad_symm(vn, phi, cache):
    """surface_grad_symm!(tau::VectorData, v::Edges) (:643)"""
    _need_vector(cache, "surface_grad_symm")
    _expect(vn, VectorData, "surface_grad_symm")
    _expect(phi, Edges, "surface_grad_symm")
    L.check(cache._lib.ilm_vsurface_grad(cache._plan, 1, _ptr(phi.data), _ptr(vn.data)))
    return vn


def surface_curl(a, b, cache):
    """surface_curl! in its four forms (:357, :388, :412, :443)"""
    if isinstance(b, VectorData):
        _need_vector(cache, "surface_curl")
        _expect_nodes(a, Dual, "surface_curl")
        L.check(cache._lib.ilm_vsurface_curl_s2n(cache._plan, _ptr(b.data), _ptr(a.data)))
        return a
    if isinstance(a, VectorData):
        _need_vector(cache, "surface_curl")
        _expect_nodes(b, Dual, "surface_curl")
        L.check(cache._lib.ilm_vsurface_curl_n2s(cache._plan, _ptr(b.data), _ptr(a.data)))
        return a
    return _s_surface_curl(a, b, cache)


def _mask_layout(w, cache):
    """Layout code of grid data `w` for mask!/complementary_mask! (the methods of
    _scalar_mask_product! / _vector_mask_product!, src/surface_operators.jl:880-923)."""
    if isinstance(w, Nodes):
        return L.NODES_PRIMAL if w.celltype == Primal else L.NODES_DUAL
    if isinstance(w, XEdges):
        return L.XEDGES
    if isinstance(w, YEdges):
        return L.YEDGES
    if isinstance(w, Edges):
        return L.EDGES
    if isinstance(w, EdgeGradient):
        return L.EDGEGRAD
    raise MethodError(f"mask!: no method for {type(w).__name__}")


def _mask_inplace(w, cache, complementary):
    if cache.scaling != GridScaling:
        raise MethodError("mask! is only defined for GridScaling caches")
    layout = _mask_layout(w, cache)
    vec = _is_vector(cache)
    if vec and layout in (L.XEDGES, L.YEDGES):
        raise MethodError("mask!: no method for a single edge component on a vector cache")
    if not vec and layout == L.EDGEGRAD:
        raise MethodError("mask!: no method for EdgeGradient on a scalar cache")
    L.check(cache._lib.ilm_mask_product(cache._plan, L.VECTOR_CACHE if vec else L.SCALAR_CACHE, layout,
                                        1 if complementary else 0, _ptr(w.data)))
    return w


def _mask_cache(w, surface, g, **kwargs):
    """_mask_cache (src/surface_operators.jl:846-848): a scalar cache for scalar grid data, a vector cache for Edges /
    EdgeGradient.  kwargs: `parent=` (a cache on the same grid whose Laplacian is shared) or `lgf_table=`."""
    cls = SurfaceVectorCache if isinstance(w, (Edges, EdgeGradient)) else SurfaceScalarCache
    return cls(surface, g, device=_is_torch(w.data), **kwargs)


def mask(*args, **kwargs):
    """mask(cache) -> new grid data with 1 inside / 0 outside (src/surface_operators.jl:737): Nodes{Primal}
    for a scalar cache, Edges for a vector cache.  mask(w, cache) = mask!(w, cache) (:788-800): w is
    multiplied in place by the mask, averaged onto w's layout where needed.
    mask(w, surface, g) = mask!(w, surface, grid) (:826-829): the same with a temporary cache of the shape."""
    if len(args) == 3:
        return _mask_inplace(args[0], _mask_cache(args[0], args[1], args[2], **kwargs), False)
    if len(args) == 2:
        return _mask_inplace(args[0], args[1], False)
    (cache,) = args
    if _is_vector(cache):
        if cache.scaling != GridScaling:
            raise MethodError("mask is only defined for GridScaling caches")
        m = cache.zeros_grid()
        L.check(cache._lib.ilm_mask_edges(cache._plan, _ptr(m.data)))
        return m
    return _s_mask(cache)


def complementary_mask(*args, **kwargs):
    """complementary_mask(cache) (:749) / complementary_mask!(w, cache) (:811-823) /
    complementary_mask!(w, surface, grid) (:837-840)."""
    if len(args) == 3:
        return _mask_inplace(args[0], _mask_cache(args[0], args[1], args[2], **kwargs), True)
    if len(args) == 2:
        return _mask_inplace(args[0], args[1], True)
    (cache,) = args
    m = mask(cache)
    if _is_torch(m.data):
        m.data.neg_().add_(1.0)
    else:
        m.data[...] = 1.0 - m.data
    return m


def _vmatrix(cache, which, scale, cols):
    M = 2 * cache.N
    c0, c1 = (0, M) if cols is None else cols
    if cache.device:
        import torch
        buf = torch.zeros(M * (c1 - c0), dtype=torch.float64, device="cuda")
    else:
        buf = _host_zeros(M * (c1 - c0))
    L.check(cache._lib.ilm_create_schur_vector(cache._plan, which, float(scale), int(c0), int(c1), _ptr(buf)))
    return _as_matrix(buf, M, c1 - c0)


def create_RTLinvR(cache, scale=1.0, cols=None):
    """create_RTLinvR (src/matrix_operators.jl:9-30): N x N (scalar cache) or 2N x 2N (vector cache)."""
    return _vmatrix(cache, 0, scale, cols) if _is_vector(cache) else _s_create_RTLinvR(cache, scale, cols)


def create_CLinvCT(cache, scale=1.0, cols=None):
    """create_CLinvCT (:40-61) on the cache's primary point data type."""
    return _vmatrix(cache, 1, scale, cols) if _is_vector(cache) else _s_create_CLinvCT(cache, scale, cols)


def create_CLinvCT_scalar(cache, scale=1.0, cols=None):
    """create_CLinvCT_scalar (:71-92): ScalarData form, vector caches only."""
    _need_vector(cache, "create_CLinvCT_scalar")
    return _s_create_CLinvCT(cache, scale, cols)


def create_CL2invCT(cache, scale=1.0, cols=None):
    """create_CL2invCT (:102-125): -C_s L^-2 C_s^T."""
    _need_vector(cache, "create_CL2invCT")
    return _vmatrix(cache, 2, scale, cols)


def create_GLinvD(cache, scale=1.0, cols=None):
    """create_GLinvD (:135-155)."""
    return _vmatrix(cache, 3, scale, cols) if _is_vector(cache) else _s_create_GLinvD(cache, scale, cols)


def create_GLinvD_symm(cache, scale=1.0, cols=None):
    """create_GLinvD_symm (:165-185)."""
    _need_vector(cache, "create_GLinvD_symm")
    return _vmatrix(cache, 4, scale, cols)


def create_nRTRn(cache, scale=1.0):
    """create_nRTRn (:225-244)."""
    if not _is_vector(cache):
        return _s_create_nRTRn(cache, scale)
    M = 2 * cache.N
    buf = _matrix(cache, 4 * cache.N) if cache.N else _matrix(cache, 0)
    L.check(cache._lib.ilm_create_nRTRn_vector(cache._plan, float(scale), _ptr(buf)))
    return _as_matrix(buf, M, M)


def create_RTLinvR_direct(cache, scale=1.0, cols=None):
    """The matrix of create_RTLinvR from the direct LGF-table identity:
    no transform is applied.  Cross-check and optional fast builder for scalar caches."""
    if _is_vector(cache):
        raise MethodError("create_RTLinvR_direct: scalar caches only")
    c0, c1 = (0, cache.N) if cols is None else cols
    buf = _matrix(cache, c1 - c0)
    L.check(cache._lib.ilm_create_RTLinvR_direct(cache._plan, float(scale), int(c0), int(c1), _ptr(buf)))
    return _as_matrix(buf, cache.N, c1 - c0)


def neumann_poisson(cache, vnplus, vnminus=None, S=None):
    """The Neumann Poisson solve of test/literate/neumann.jl:101-142 on the B200 path (config C2):
    S = create_CLinvCT(cache); returns (f, df, s, ds) = potential, its surface jump, streamfunction
    and its jump.  Host-mode or device-mode containers follow the cache."""
    N = cache.N
    vnplus = np.asarray(vnplus, dtype=np.float64)
    vnminus = np.zeros(N) if vnminus is None else np.asarray(vnminus, dtype=np.float64)
    if S is None:
        S = create_CLinvCT(cache)
    lu = S if isinstance(S, LU) else LU(S)                  # an LU object is reused as is
    dvn = cache.zeros_surface().set(vnplus - vnminus)
    vn = 0.5 * (vnplus + vnminus)
    fstar = cache.zeros_grid()
    regularize(fstar, dvn, cache)
    inverse_laplacian(fstar, cache)
    df = cache.zeros_surface()
    surface_grad(df, fstar, cache)
    df.set(_neg_solve(lu, vn, df, cache))
    f = cache.zeros_grid()
    surface_divergence(f, df, cache)
    inverse_laplacian(f, cache)
    _iadd(f, fstar)
    # streamfunction
    sstar = cache.zeros_gridcurl()
    surface_curl(sstar, df, cache)
    ds = cache.zeros_surface()
    surface_grad_cross(ds, fstar, cache)
    sol = lu.solve(ds.data)
    _assign(ds, sol)
    s = cache.zeros_gridcurl()
    surface_curl_cross(s, ds, cache)
    _isub(s, sstar)
    _iscale(s, -1.0)
    inverse_laplacian(s, cache)
    return f, df, s, ds


def _to_host(x):
    return x.detach().cpu().numpy() if _is_torch(x) else np.asarray(x)


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _neg_solve(lu, vn, df, cache):
    """-(S \\ (vn - df)) as a host array."""
    rhs = (_dev(vn) - df.data) if cache.device else (vn - df.data)
    return -_to_host(lu.solve(rhs))


def _assign(d, values):
    if _is_torch(d.data):
        d.data.copy_(values if _is_torch(values) else _dev(values))
    else:
        d.data[...] = _to_host(values)


def _iadd(a, b):
    if _is_torch(a.data):
        a.data.add_(b.data)
    else:
        a.data += b.data


def _isub(a, b):
    if _is_torch(a.data):
        a.data.sub_(b.data)
    else:
        a.data -= b.data


def _iscale(a, c):
    if _is_torch(a.data):
        a.data.mul_(c)
    else:
        a.data *= c


def create_RTHR(cache, kernel_id, scale=1.0, cols=None):
    """-scale * E H R with H the convolution kernel `kernel_id` registered with
    `cache.add_kernel` (e.g. the integrating factor exp(L a), `lgf.intfact_table`): the Schur
    complement `S_i = -B2 H_i B1^T` of an IF-HERK stage (src/timemarching.jl:86-107 through
    ConstrainedSystems).  kernel_id = 0 is create_RTLinvR."""
    if _is_vector(cache):
        raise MethodError("create_RTHR: scalar caches only")
    c0, c1 = (0, cache.N) if cols is None else cols
    buf = _matrix(cache, c1 - c0)
    L.check(cache._lib.ilm_create_schur_kernel(cache._plan, L.RTLINVR, int(kernel_id), float(scale), int(c0), int(c1), _ptr(buf)))
    return _as_matrix(buf, cache.N, c1 - c0)


def create_RTHR_direct(cache, table, c0=0.0, factor=1.0, scale=1.0, cols=None):
    """-scale/factor * E (T - c0) R from the kernel table T itself (no transform): the fast builder
    of an IF-HERK stage complement for a compact table such as `lgf.intfact_table(a, 32)`; what makes
    the per-step operator refresh of a moving body affordable."""
    if _is_vector(cache):
        raise MethodError("create_RTHR_direct: scalar caches only")
    table = np.asfortranarray(table, dtype=np.float64)
    if table.ndim != 2 or table.shape[0] != table.shape[1]:
        raise DimensionMismatch("create_RTHR_direct: square kernel table expected")
    c0_, c1_ = (0, cache.N) if cols is None else cols
    buf = _matrix(cache, c1_ - c0_)
    L.check(cache._lib.ilm_create_schur_direct_kernel(cache._plan, _ptr(table), table.shape[0], float(c0), float(factor),
                                                      float(scale), int(c0_), int(c1_), _ptr(buf)))
    return _as_matrix(buf, cache.N, c1_ - c0_)


def convolve(w, cache, kernel_id):
    """w <- K * w for a kernel registered with `cache.add_kernel`: `exp(L, a) * w` (plan_intfact, the
    integrating factor of src/timemarching.jl:93,104) or `implicit_operator(L, a) \\ w`
    (src/grid_operators.jl:182-184).  In place; Nodes, XEdges / YEdges or Edges."""
    if isinstance(w, Edges):
        layout = L.EDGES
    elif isinstance(w, (Nodes, XEdges, YEdges)):
        layout = w.layout
    else:
        raise MethodError(f"convolve: no method for {type(w).__name__}")
    L.check(cache._lib.ilm_convolve(cache._plan, int(kernel_id), layout, _ptr(w.data)))
    return w


def stokes_flow(cache, vplus, vminus=None, S=None, Ss=None):
    """Stokes flow past a body with prescribed surface velocities, `solve(prob::StokesFlowProblem, sys)` of
    test/literate/stokes.jl:98-166 on the B200 path (BASELINE config C5a: vector-data / Rf problems).
    vplus, vminus: (2N,) exterior / interior surface velocities [u; v].  Returns (v, s, sigma, S, Ss):
    velocity (Edges), streamfunction (Nodes{Dual}), traction (VectorData) and the two factored-once
    matrices S = create_CL2invCT, Ss = create_CLinvCT_scalar (reusable for other boundary data).
    The final traction filter `C^2 * sigma` (:163) is left to the caller (create_surface_filter)."""
    _need_vector(cache, "stokes_flow")
    N = cache.N
    vplus = np.asarray(vplus, dtype=np.float64)
    vminus = np.zeros(2 * N) if vminus is None else np.asarray(vminus, dtype=np.float64)
    nx, ny = cache.normals()
    dv_h = vplus - vminus                                       # prescribed_surface_jump!
    dv = cache.zeros_surface().set(dv_h)
    v = cache.zeros_grid()
    sstar = cache.zeros_gridcurl()
    surface_divergence_symm(v, dv, cache)                       # psi*: -L^-2 curl(D_s^symm dv)
    curl(sstar, v, cache)
    _iscale(sstar, -1.0)
    inverse_laplacian(sstar, cache)
    inverse_laplacian(sstar, cache)
    dvn = cache.zeros_surfacescalar().set(nx * dv_h[:N] + ny * dv_h[N:])      # pointwise_dot!(dvn, nrm, dv)
    phi = cache.zeros_griddiv()
    regularize(phi, dvn, cache)
    inverse_laplacian(phi, cache)
    vphi = cache.zeros_grid()
    grad(vphi, phi, cache)
    vb = cache.zeros_surface()
    interpolate(vb, vphi, cache)
    vprime = cache.zeros_surface().set(0.5 * (vplus + vminus))  # prescribed_surface_average!
    _isub(vprime, vb)
    curl(v, sstar, cache)
    interpolate(vb, v, cache)
    _isub(vprime, vb)                                           # spurious slip
    if S is None:
        S = LU(create_CL2invCT(cache))
    sigma = cache.zeros_surface()
    _assign(sigma, S.solve(vprime.data))
    s = cache.zeros_gridcurl()
    surface_curl(s, sigma, cache)                               # correction streamfunction
    _iscale(s, -1.0)
    inverse_laplacian(s, cache)
    inverse_laplacian(s, cache)
    _iadd(s, sstar)
    curl(v, s, cache)
    _iadd(v, vphi)
    ds = cache.zeros_surfacescalar()                            # streamfunction equivalent of the potential
    surface_grad_cross(ds, phi, cache)
    if Ss is None:
        Ss = LU(create_CLinvCT_scalar(cache))
    _assign(ds, Ss.solve(ds.data))
    surface_curl_cross(sstar, ds, cache)
    _iscale(sstar, -1.0)
    inverse_laplacian(sstar, cache)
    _iadd(s, sstar)
    return v, s, sigma, S, Ss


def convective_derivative(out, v, *args):
    """convective_derivative!(udp, v, p, cache) = v . grad p on Nodes{Primal} (src/grid_operators.jl:258-264)
    or on Nodes{Dual} (:272-288);
    convective_derivative!(vdu, v, u, cache) = (v . grad) u on Edges (:290-299);
    convective_derivative!(udu, u, cache) = (u . grad) u (:308-316).  Divided by dx for GridScaling.
    No ConvectiveDerivativeCache argument: the fused kernels need no temporaries."""
    if len(args) == 1:
        (cache,), q = args, v
    elif len(args) == 2:
        q, cache = args
    else:
        raise MethodError("convective_derivative: expected (out, v, cache) or (out, v, q, cache)")
    _expect(v, Edges, "convective_derivative")
    if isinstance(q, Nodes):
        _expect_nodes(out, q.celltype, "convective_derivative")
        fn = cache._lib.ilm_convective_derivative_scalar if q.celltype == Primal else cache._lib.ilm_convective_derivative_dual
        L.check(fn(cache._plan, _ptr(v.data), _ptr(q.data), _ptr(out.data)))
        return out
    _expect(q, Edges, "convective_derivative")
    _expect(out, Edges, "convective_derivative")
    if out is v or out is q:
        raise MethodError("convective_derivative: the result must not alias an input")
    L.check(cache._lib.ilm_convective_derivative_vector(cache._plan, _ptr(v.data), _ptr(q.data), _ptr(out.data)))
    return out


def w_cross_v(vw, w, v, cache):
    """w_cross_v!(vw::Edges, w::Nodes{Dual}, v::Edges, cache) (src/grid_operators.jl:404-434): the
    rotational form of the convective term; not scaled by the grid spacing."""
    _expect(vw, Edges, "w_cross_v")
    _expect_nodes(w, Dual, "w_cross_v")
    _expect(v, Edges, "w_cross_v")
    if vw is v:
        raise MethodError("w_cross_v: the result must not alias the velocity")
    L.check(cache._lib.ilm_w_cross_v(cache._plan, _ptr(w.data), _ptr(v.data), _ptr(vw.data)))
    return vw



# --------------------------------------------------------------------------
# Every operator validates its containers against the cache before the C call: the library copies
# `ilm_layout_size(plan, layout)` doubles through the raw pointers, so a container built on another grid (or
# point data of another length) would be a host heap overrun / device out-of-bounds access.  The reference gets
# this from the NX, NY type parameters (MethodError) and the length asserts of src/tools.jl:85,102.
# --------------------------------------------------------------------------
def _expected_len(d, cache):
    g, N = cache.g, cache.N
    if isinstance(d, (Nodes, _EdgeComponent)):
        mx, my = g.layout_shape(d.layout)
        return mx * my
    if isinstance(d, Edges):
        (ux, uy), (vx, vy) = g.layout_shape(L.XEDGES), g.layout_shape(L.YEDGES)
        return ux * uy + vx * vy
    if isinstance(d, EdgeGradient):
        (px, py), (dx_, dy_) = g.layout_shape(L.NODES_PRIMAL), g.layout_shape(L.NODES_DUAL)
        return 2 * px * py + 2 * dx_ * dy_
    if isinstance(d, TensorData):
        return 4 * N
    if isinstance(d, VectorData):
        return 2 * N
    if isinstance(d, ScalarData):
        return N
    return None


def _check_containers(what, args):
    cache = next((a for a in reversed(args) if isinstance(a, SurfaceScalarCache)), None)
    if cache is None:
        return
    for a in args:
        if isinstance(a, _Data):
            n = _expected_len(a, cache)
            if n is not None and len(a) != n:
                raise DimensionMismatch(f"{what}: {type(a).__name__} of length {len(a)} does not belong to this cache "
                                        f"(grid {cache.g.NX} x {cache.g.NY}, {cache.N} points: expected {n})")


_NOT_LINEAR = {"convective_derivative", "w_cross_v"}


def _complex_call(fn, args, kwargs):
    """An operator on complex containers (dtype = ComplexF64 caches): the operators of the path are real-linear, so
    the real and the imaginary parts go through the library one after the other and are merged back."""
    import copy
    if fn.__name__ in _NOT_LINEAR:
        raise MethodError(f"{fn.__name__}: not defined for complex data (the operator is bilinear)")
    data = [a for a in args if isinstance(a, _Data)]
    if not all(a.is_complex for a in data):
        raise MethodError(f"{fn.__name__}: real and complex containers cannot be mixed")
    parts = []
    for part in ("real", "imag"):
        clones = {}
        for a in data:
            if id(a) in clones:
                continue
            c = copy.copy(a)
            src = getattr(a.data, part)
            c.data = src.contiguous().clone() if _is_torch(src) else np.ascontiguousarray(src).copy()
            clones[id(a)] = c
        fn(*[clones[id(a)] if isinstance(a, _Data) else a for a in args], **kwargs)
        parts.append(clones)
    for a in data:
        re, im = parts[0][id(a)].data, parts[1][id(a)].data
        if _is_torch(a.data):
            import torch
            a.data.copy_(torch.complex(re, im))
        else:
            a.data[...] = re + 1j * im
        for clone in (parts[0][id(a)], parts[1][id(a)]):
            clone.data = np.zeros(0)                   # the clones must not recycle the originals' buffers
    return args[0]


def _checked(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        _check_containers(fn.__name__, args)
        if any(isinstance(a, _Data) and a.is_complex for a in args):
            return _complex_call(fn, args, kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _wrap_operators():
    import inspect
    g = globals()
    for name, fn in list(g.items()):
        if name.startswith("_") or not inspect.isfunction(fn) or fn.__module__ != __name__:
            continue
        params = inspect.signature(fn).parameters
        if "cache" in params or any(p.kind is inspect.Parameter.VAR_POSITIONAL for p in params.values()):
            g[name] = _checked(fn)


_wrap_operators()


def _wrap_zeros(cls):
    """zeros_* honour the cache's element type; similar_* are the reference's names (src/cache.jl:395-440)."""
    import functools
    for name, fn in list(vars(cls).items()):
        if name.startswith("zeros_") and callable(fn):
            def make(f):
                @functools.wraps(f)
                def w(self, *a, **k):
                    return self._cplx(f(self, *a, **k))
                return w
            setattr(cls, name, make(fn))
            setattr(cls, "similar_" + name[len("zeros_"):], getattr(cls, name))


_wrap_zeros(SurfaceScalarCache)
_wrap_zeros(SurfaceVectorCache)

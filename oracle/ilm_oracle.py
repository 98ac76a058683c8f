"""CPU oracle for the immersed-layer operator hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy/scipy restatement of the arithmetic that
JuliaIBPM/ImmersedLayers.jl v0.5.6 performs on its hot path (regularize! /
interpolate!, the staggered stencils, the lattice-Green's-function inverse
Laplacian, the Schur-complement builders and the Dirichlet block-LU solve).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it; the product (``immersedlayers.jl_b200``)
never does.

PARITY PINNING.  The path's inner arithmetic lives in un-vendored Julia
dependencies of the reference (CartesianGrids.jl, compat "0.1.29, 0.2",
/root/reference/Project.toml:8,20; no Manifest) and there is no Julia
toolchain here, so the reference cannot be executed.  The restatement follows
SURVEY.md Appendix A (published algorithm of CartesianGrids: DDF formulas,
staggered index ranges, LGF integral, CircularConvolution) and the
*composition* code that IS in the reference (file:line cited per function).
It is pinned against everything the reference holds for this path:
  * the adjoint / partition-of-unity identities of test/tools.jl:107,112,119-120,
  * the executed-notebook golden values of SURVEY.md Appendix B
    (examples/caches.ipynb, examples/Layers.ipynb),
  * exact LGF values G(1,0)=1/4, G(1,1)=1/pi, G(2,0)=1-2/pi and L*G=delta,
  * physics checks of test/surface_ops.jl (mask integral, operator norms),
  * two FULL-PROBLEM results printed by the reference's executed notebooks, reproduced to 2e-14
    (tests/test_golden.py): the temperature at (-0.9, 0) after 51 and 54 IF-HERK steps of
    examples/heatconduction.ipynb (cells 58, 62: DDF tables, surface operators, plan_intfact and the
    un-vendored ConstrainedSystems integrator) and the added mass 0.22649277527914002 of a circle in a
    square box from examples/neumann.ipynb (cell 37: LGF inverse Laplacian, create_CLinvCT, surface
    grad / divergence / curl, body lists, the block-LU Neumann solve).
What stays "parity unpinned" (no reference output fixes it): the far-field
constant c0 (it multiplies the net source, which is zero in both pinned problems), the surface filter
and the tensor-component order of the vector cache.

Conventions: size(g)=(NX,NY) dual cells incl. ghosts; arrays
are indexed a[i-1, j-1] for the 1-based Julia index (i, j), x first; flattening
with order='F' gives Julia's memory layout.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.linalg
import scipy.sparse as sp

# --------------------------------------------------------------------------
# layouts
# --------------------------------------------------------------------------
PRIMAL, DUAL, XEDGE, YEDGE = "primal", "dual", "xedge", "yedge"
KINDS = (PRIMAL, DUAL, XEDGE, YEDGE)


def field_shape(kind, NX, NY):
    """Element counts per layout: Nodes{Primal} (NX-1,NY-1); Nodes{Dual}
    (NX,NY); Edges{Primal}.u (NX,NY-1); Edges{Primal}.v (NX-1,NY)."""
    return {PRIMAL: (NX - 1, NY - 1), DUAL: (NX, NY),
            XEDGE: (NX, NY - 1), YEDGE: (NX - 1, NY)}[kind]


def field_shift(kind):
    """Index-space offset of entry (i,j): position = (i - sx, j - sy)."""
    return {PRIMAL: (0.0, 0.0), DUAL: (0.5, 0.5),
            XEDGE: (0.5, 0.0), YEDGE: (0.0, 0.5)}[kind]


class Grid:
    """PhysicalGrid surrogate: (NX, NY, dx, I0) passed explicitly
    (upstream auto-tunes the size, never re-derive it)."""

    def __init__(self, NX, NY, dx, I0):
        self.NX, self.NY, self.dx = int(NX), int(NY), float(dx)
        self.I0 = (int(I0[0]), int(I0[1]))

    def zeros(self, kind):
        return np.zeros(field_shape(kind, self.NX, self.NY), order="F")

    def coords(self, kind):
        mx, my = field_shape(kind, self.NX, self.NY)
        sx, sy = field_shift(kind)
        x = (np.arange(1, mx + 1) - sx - self.I0[0]) * self.dx
        y = (np.arange(1, my + 1) - sy - self.I0[1]) * self.dx
        return x, y


# --------------------------------------------------------------------------
# discrete delta functions.  The operation order below is the
# contract shared with csrc/ilm_ddf.h so that tables are bit-exact.
# --------------------------------------------------------------------------
_C1 = 0.40454998233983938      # 17/48 + sqrt(3) pi/108
_C2 = 1.0954500176601606       # 55/48 - sqrt(3) pi/108
_S12 = 0.14433756729740644     # sqrt(3)/12
_S2 = 0.86602540378443865      # sqrt(3)/2
_S36 = 0.048112522432468814    # sqrt(3)/36
_C1312 = 1.0833333333333333    # 13/12
_C148 = 0.020833333333333332   # 1/48

# fdlibm e_asin.c constants (Sun Microsystems, the algorithm Julia's Base.asin
# also implements)
_pio2_hi = 1.57079632679489655800e+00
_pio2_lo = 6.12323399573676603587e-17
_pio4_hi = 7.85398163397448278999e-01
_pS = (1.66666666666666657415e-01, -3.25565818622400915405e-01,
       2.01212532134862925881e-01, -4.00555345006794114027e-02,
       7.91534994289814532176e-04, 3.47933107596021167570e-05)
_qS = (-2.40339491173441421878e+00, 2.02094576023350569471e+00,
       -6.88283971605453293030e-01, 7.70381505559019352791e-02)


def asin_fdlibm(x):
    """asin restated from the published fdlibm algorithm with a fixed
    operation order (no FMA), valid for |x| <= 1."""
    x = np.asarray(x, dtype=np.float64)
    ax = np.abs(x)
    out = np.empty_like(x)

    def pq(t):
        p = t * (_pS[0] + t * (_pS[1] + t * (_pS[2] + t * (_pS[3] + t * (_pS[4] + t * _pS[5])))))
        q = 1.0 + t * (_qS[0] + t * (_qS[1] + t * (_qS[2] + t * _qS[3])))
        return p, q

    small = ax < 0.5
    tiny = ax < 2.0 ** -27
    xs = x[small]
    t = xs * xs
    p, q = pq(t)
    w = p / q
    res = xs + xs * w
    res = np.where(tiny[small], xs, res)
    out[small] = res

    big = ~small
    xb = ax[big]
    w = 1.0 - xb
    t = w * 0.5
    p, q = pq(t)
    s = np.sqrt(t)
    # |x| >= 0.975
    w1 = p / q
    t1 = _pio2_hi - (2.0 * (s + s * w1) - _pio2_lo)
    # 0.5 <= |x| < 0.975
    wl = (s.view(np.uint64) & np.uint64(0xFFFFFFFF00000000)).view(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (t - wl * wl) / (s + wl)
    r = p / q
    p2 = 2.0 * s * r - (_pio2_lo - 2.0 * c)
    q2 = _pio4_hi - 2.0 * wl
    t2 = _pio4_hi - (p2 - q2)
    # fdlibm tests the high word: ix >= 0x3FEF3333
    hi = (xb.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    tb = np.where(hi >= 0x3FEF3333, t1, t2)
    tb = np.where(xb >= 1.0, xb * _pio2_hi + xb * _pio2_lo, tb)
    out[big] = np.where(x[big] > 0, tb, -tb)
    return out


def ddf_yang3(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    out = np.zeros_like(r)
    a = r <= 1.0
    ra = r[a]
    rr = ra * ra
    s = np.sqrt(((-12.0 * rr) + 12.0 * ra) + 1.0)
    t3 = ((1.0 - 2.0 * ra) * 0.0625) * s
    t4 = _S12 * asin_fdlibm(_S2 * (2.0 * ra - 1.0))
    out[a] = (((_C1 + ra * 0.25) - rr * 0.25) + t3) - t4
    b = (r > 1.0) & (r < 2.0)
    rb = r[b]
    rr = rb * rb
    s = np.sqrt(((-12.0 * rr) + 36.0 * rb) - 23.0)
    t3 = ((2.0 * rb - 3.0) * _C148) * s
    t4 = _S36 * asin_fdlibm(_S2 * (2.0 * rb - 3.0))
    out[b] = (((_C2 - _C1312 * rb) + rr * 0.25) + t3) + t4
    return out


def ddf_m3(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    out = np.zeros_like(r)
    a = r <= 0.5
    out[a] = 0.75 - r[a] * r[a]
    b = (r > 0.5) & (r < 1.5)
    d = 1.5 - r[b]
    out[b] = 0.5 * (d * d)
    return out


def ddf_roma(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    out = np.zeros_like(r)
    a = r <= 0.5
    out[a] = (1.0 + np.sqrt(1.0 - 3.0 * (r[a] * r[a]))) / 3.0
    b = (r > 0.5) & (r < 1.5)
    d = 1.0 - r[b]
    out[b] = ((5.0 - 3.0 * r[b]) - np.sqrt(1.0 - 3.0 * (d * d))) / 6.0
    return out


def ddf_m4prime(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    out = np.zeros_like(r)
    a = r <= 1.0
    rr = r[a] * r[a]
    out[a] = (1.0 - 2.5 * rr) + 1.5 * (rr * r[a])
    b = (r > 1.0) & (r < 2.0)
    d = 2.0 - r[b]
    out[b] = (0.5 * (d * d)) * (1.0 - r[b])
    return out


def ddf_witchhat(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    return np.where(r < 1.0, 1.0 - r, 0.0)


# name -> (function, support radius rho, window width W)
DDFS = {
    "yang3": (ddf_yang3, 2.0, 4),
    "m3": (ddf_m3, 1.5, 3),
    "roma": (ddf_roma, 1.5, 3),
    "m4prime": (ddf_m4prime, 2.0, 4),
    "witchhat": (ddf_witchhat, 1.0, 2),
}


# --------------------------------------------------------------------------
# Regularize / R / E tables (reference call sites
# src/cache.jl:305-314 (weights), :328-347 (matrices))
# --------------------------------------------------------------------------
GRID_SCALING, INDEX_SCALING = "grid", "index"


class Table:
    """Per-point window table for one target layout.
    i0,j0: 0-based array index of the first window entry (may be out of range,
    entries outside the field are masked); wR (N,W,W) regularization weights
    indexed [k, a(x), b(y)]; wE same for interpolation; valid mask."""

    def __init__(self, kind, shape, i0, j0, wR, wE, valid):
        self.kind, self.shape = kind, shape
        self.i0, self.j0, self.wR, self.wE, self.valid = i0, j0, wR, wE, valid
        self.N, self.W = wR.shape[0], wR.shape[1]

    def linear_index(self):
        """0-based column-major linear index (N,W,W); -1 where masked."""
        mx, _ = self.shape
        a = np.arange(self.W)
        ii = self.i0[:, None, None] + a[None, :, None]
        jj = self.j0[:, None, None] + a[None, None, :]
        lin = ii + mx * jj
        return np.where(self.valid, lin, -1)


def scaled_points(grid, x, y):
    """xs = x/dx + I0 (A.2)."""
    return np.asarray(x) / grid.dx + grid.I0[0], np.asarray(y) / grid.dx + grid.I0[1]


def point_weights(grid, ds, scaling):
    """GridScaling: w = ds/dx^2 (src/cache.jl:306-307 passes weights=a.data and
    upstream divides by dx*dx); IndexScaling: w = 1 (issymmetric, :309-310)."""
    ds = np.asarray(ds, dtype=np.float64)
    if scaling == GRID_SCALING:
        return ds / (grid.dx * grid.dx)
    return np.ones_like(ds)


def build_table(grid, x, y, ds, kind, ddf="yang3", scaling=GRID_SCALING):
    fn, rho, W = DDFS[ddf]
    xs, ys = scaled_points(grid, x, y)
    wgt = point_weights(grid, ds, scaling)
    sx, sy = field_shift(kind)
    mx, my = field_shape(kind, grid.NX, grid.NY)
    N = xs.shape[0]
    # first 1-based grid index of the window
    ib = np.floor((xs + sx) - rho).astype(np.int64) + 1
    jb = np.floor((ys + sy) - rho).astype(np.int64) + 1
    a = np.arange(W)
    ri = ((ib[:, None] + a[None, :]) - sx) - xs[:, None]
    rj = ((jb[:, None] + a[None, :]) - sy) - ys[:, None]
    px = fn(ri)
    py = fn(rj)
    wE = px[:, :, None] * py[:, None, :]
    wR = wE * wgt[:, None, None]
    i1 = ib[:, None, None] + a[None, :, None]
    j1 = jb[:, None, None] + a[None, None, :]
    valid = (i1 >= 1) & (i1 <= mx) & (j1 >= 1) & (j1 <= my)
    valid = np.broadcast_to(valid, (N, W, W)).copy()
    wR = np.where(valid, wR, 0.0)
    wE = np.where(valid, wE, 0.0)
    if scaling == INDEX_SCALING:
        wE = wR.copy()
    return Table(kind, (mx, my), ib - 1, jb - 1, wR, wE, valid)


def regularize(tab, f):
    """s = R f (src/surface_operators.jl:38-41: fill!(s,0); mul!(s,Rc,f)).
    CSC mat-vec order: columns (points) ascending, y[row] += a*x[col], the
    product rounded before the add."""
    mx, my = tab.shape
    out = np.zeros(mx * my)
    lin = tab.linear_index()
    # order (k, b, a): rows ascending inside a column
    lin_o = np.transpose(lin, (0, 2, 1)).reshape(-1)
    val = np.transpose(tab.wR * np.asarray(f)[:, None, None], (0, 2, 1)).reshape(-1)
    m = lin_o >= 0
    np.add.at(out, lin_o[m], val[m])
    return out.reshape((mx, my), order="F")


def interpolate(tab, s):
    """f = E s (src/surface_operators.jl:80-83).  CSC mat-vec over grid
    columns in ascending linear index: per point the sum runs b (y) outer,
    a (x) inner."""
    s = np.asarray(s)
    N, W = tab.N, tab.W
    f = np.zeros(N)
    mx, my = tab.shape
    for b in range(W):
        for a in range(W):
            v = tab.valid[:, a, b]
            ii = np.clip(tab.i0 + a, 0, mx - 1)
            jj = np.clip(tab.j0 + b, 0, my - 1)
            term = tab.wE[:, a, b] * s[ii, jj]
            f = f + np.where(v, term, 0.0)
    return f


def R_matrix(tab):
    """scipy CSC regularization matrix (grid x N), the structure upstream uses."""
    lin = tab.linear_index()
    k = np.broadcast_to(np.arange(tab.N)[:, None, None], lin.shape)
    m = lin >= 0
    mx, my = tab.shape
    return sp.csc_matrix((tab.wR[m], (lin[m], k[m])), shape=(mx * my, tab.N))


def E_matrix(tab):
    lin = tab.linear_index()
    k = np.broadcast_to(np.arange(tab.N)[:, None, None], lin.shape)
    m = lin >= 0
    mx, my = tab.shape
    return sp.csc_matrix((tab.wE[m], (k[m], lin[m])), shape=(tab.N, mx * my))


# --------------------------------------------------------------------------
# staggered stencils, unscaled.  Outputs are zero outside the
# stated ranges (the reference zero-fills first: src/grid_operators.jl:26,46,
# 62,80,98).  Python slices below are the 1-based ranges shifted by one.
# --------------------------------------------------------------------------
def curl_n2e(grid, s):
    """curl!(Edges{Primal} <- Nodes{Dual})."""
    NX, NY = grid.NX, grid.NY
    u = np.zeros((NX, NY - 1), order="F")
    v = np.zeros((NX - 1, NY), order="F")
    u[:, :] = s[:, 1:] - s[:, :-1]
    v[:, :] = s[:-1, :] - s[1:, :]
    return u, v


def curl_e2n(grid, u, v):
    """curl!(Nodes{Dual} <- Edges{Primal}): w[x,y]=u[x,y-1]-u[x,y]-v[x-1,y]+v[x,y],
    x in 2:NX-1, y in 2:NY-1."""
    NX, NY = grid.NX, grid.NY
    w = np.zeros((NX, NY), order="F")
    w[1:NX - 1, 1:NY - 1] = ((u[1:NX - 1, 0:NY - 2] - u[1:NX - 1, 1:NY - 1])
                             - v[0:NX - 2, 1:NY - 1]) + v[1:NX - 1, 1:NY - 1]
    return w


def divergence_e2n(grid, u, v):
    """divergence!(Nodes{Primal} <- Edges{Primal}): p[x,y]=-u[x,y]+u[x+1,y]-v[x,y]+v[x,y+1]."""
    NX, NY = grid.NX, grid.NY
    p = np.zeros((NX - 1, NY - 1), order="F")
    p[:, :] = ((-u[0:NX - 1, :] + u[1:NX, :]) - v[:, 0:NY - 1]) + v[:, 1:NY]
    return p


def grad_n2e(grid, p):
    """grad!(Edges{Primal} <- Nodes{Primal}): u[x,y]=p[x,y]-p[x-1,y] (x 2:NX-1),
    v[x,y]=p[x,y]-p[x,y-1] (y 2:NY-1)."""
    NX, NY = grid.NX, grid.NY
    u = np.zeros((NX, NY - 1), order="F")
    v = np.zeros((NX - 1, NY), order="F")
    u[1:NX - 1, :] = p[1:NX - 1, :] - p[0:NX - 2, :]
    v[:, 1:NY - 1] = p[:, 1:NY - 1] - p[:, 0:NY - 2]
    return u, v


def grad_e2t(grid, u, v):
    """grad!(EdgeGradient{Primal,Dual} <- Edges{Primal}); returns dudx,dudy,dvdx,dvdy."""
    NX, NY = grid.NX, grid.NY
    dudx = np.zeros((NX - 1, NY - 1), order="F")
    dvdy = np.zeros((NX - 1, NY - 1), order="F")
    dudy = np.zeros((NX, NY), order="F")
    dvdx = np.zeros((NX, NY), order="F")
    dudx[:, :] = u[1:NX, :] - u[0:NX - 1, :]
    dvdy[:, :] = v[:, 1:NY] - v[:, 0:NY - 1]
    dudy[1:NX - 1, 1:NY - 1] = u[1:NX - 1, 1:NY - 1] - u[1:NX - 1, 0:NY - 2]
    dvdx[1:NX - 1, 1:NY - 1] = v[1:NX - 1, 1:NY - 1] - v[0:NX - 2, 1:NY - 1]
    return dudx, dudy, dvdx, dvdy


def divergence_t2e(grid, dudx, dudy, dvdx, dvdy):
    """divergence!(Edges{Primal} <- EdgeGradient)."""
    NX, NY = grid.NX, grid.NY
    u = np.zeros((NX, NY - 1), order="F")
    v = np.zeros((NX - 1, NY), order="F")
    u[1:NX - 1, :] = ((dudx[1:NX - 1, :] - dudx[0:NX - 2, :])
                      + dudy[1:NX - 1, 1:NY]) - dudy[1:NX - 1, 0:NY - 1]
    v[:, 1:NY - 1] = ((dvdx[1:NX, 1:NY - 1] - dvdx[0:NX - 1, 1:NY - 1])
                      + dvdy[:, 1:NY - 1]) - dvdy[:, 0:NY - 2]
    return u, v


def grid_interpolate(grid, m, src, dst):
    """CartesianGrids grid_interpolate!(dst <- src) [UPSTREAM-RECALL, SURVEY.md A.3]: the mean of the
    nearest source entries -- 2 points along a direction in which the layouts are offset by half a
    cell, 4 points primal<->dual nodes; destination entries whose stencil leaves the source field
    stay 0 (upstream leaves its zero-initialised scratch untouched there).  Parity unpinned: the exact
    index ranges and the summation order of the 4-point mean are not fixed by any reference test."""
    NX, NY = grid.NX, grid.NY
    out = np.zeros(field_shape(dst, NX, NY))
    (sxs, sys_), (sxd, syd) = field_shift(src), field_shift(dst)

    def offs(ss, sd):                                       # source index offsets relative to the target index
        d = int(round(2 * (ss - sd)))
        return (0,) if d == 0 else ((-1, 0) if d < 0 else (0, 1))
    ox, oy = offs(sxs, sxd), offs(sys_, syd)
    mx, my = out.shape
    i = np.arange(mx)[:, None]
    j = np.arange(my)[None, :]
    ok = (i + ox[0] >= 0) & (i + ox[-1] < m.shape[0]) & (j + oy[0] >= 0) & (j + oy[-1] < m.shape[1])
    rows = []
    for b in oy:
        acc = None
        for a in ox:
            v = m[np.clip(i + a, 0, m.shape[0] - 1), np.clip(j + b, 0, m.shape[1] - 1)]
            acc = v if acc is None else acc + v
        rows.append(acc)
    tot = rows[0] if len(rows) == 1 else rows[0] + rows[1]
    out[...] = np.where(ok, (1.0 / (len(ox) * len(oy))) * tot, 0.0)
    return out


def laplacian(grid, w, kind, factor=1.0):
    """5-point stencil times L.factor on the layout's interior (A.3)."""
    mx, my = w.shape
    out = np.zeros_like(w, order="F")
    c = w[1:mx - 1, 1:my - 1]
    out[1:mx - 1, 1:my - 1] = ((((-4.0 * c + w[0:mx - 2, 1:my - 1])
                                 + w[2:mx, 1:my - 1]) + w[1:mx - 1, 0:my - 2])
                               + w[1:mx - 1, 2:my]) * factor
    return out


# --------------------------------------------------------------------------
# lattice Green's function convolution
# --------------------------------------------------------------------------
EULER_GAMMA = 0.57721566490153286


def lgf_c0(DX=1.0):
    """Far-field constant of ldiv!: (gamma + 0.5 ln 8 - ln DX)/(2 pi) (A.5)."""
    return (EULER_GAMMA + 0.5 * np.log(8.0) - np.log(DX)) / (2.0 * np.pi)


class ConvPlan:
    """CircularConvolution of upstream: kernel table G (NX,NY) mirrored to
    (2NX-1, 2NY-1), multiplier precomputed with rfft2; apply = zero-pad,
    rfft2, multiply, irfft2, crop at (NX,NY)."""

    def __init__(self, G, workers=1):
        M, N = G.shape
        self.M, self.N = M, N
        A = np.zeros((2 * M - 1, 2 * N - 1))
        A[M - 1:, N - 1:] = G
        A[:M - 1, N - 1:] = G[:0:-1, :]
        A[M - 1:, :N - 1] = G[:, :0:-1]
        A[:M - 1, :N - 1] = G[:0:-1, :0:-1]
        self.workers = workers
        self.Ghat = scipy.fft.rfft2(A, workers=workers)

    def apply(self, w):
        M, N = self.M, self.N
        mx, my = w.shape
        buf = np.zeros((2 * M - 1, 2 * N - 1))
        buf[:mx, :my] = w
        out = scipy.fft.irfft2(scipy.fft.rfft2(buf, workers=self.workers) * self.Ghat,
                               s=buf.shape, workers=self.workers)
        return np.asfortranarray(out[M - 1:M - 1 + mx, N - 1:N - 1 + my])


def inverse_laplacian(plan, w, c0, factor):
    """L\\w (A.5): out = G*w; out -= c0*sum(w); out *= 1/factor.  The iszero
    guard of src/grid_operators.jl:155 returns the input unchanged (zeros)."""
    if not np.any(w):
        return np.array(w, order="F", copy=True)
    out = plan.apply(w)
    out = out - c0 * np.sum(w)
    return out * (1.0 / factor)


def direct_convolution(G, w):
    """O(P^2) reference for small boxes: out[i,j]=sum G(|i-k|,|j-l|) w[k,l]."""
    mx, my = w.shape
    ii = np.arange(mx)
    jj = np.arange(my)
    out = np.zeros((mx, my), order="F")
    di = np.abs(ii[:, None] - ii[None, :])
    dj = np.abs(jj[:, None] - jj[None, :])
    for j in range(my):
        # out[:, j] = sum_l  (G[di, dj[j,l]] @ w[:, l])
        acc = np.zeros(mx)
        for l in range(my):
            acc += G[di, dj[j, l]] @ w[:, l]
        out[:, j] = acc
    return out


# --------------------------------------------------------------------------
# the cache ("plan") and the operator API of ImmersedLayers on it
# --------------------------------------------------------------------------
class ScalarCache:
    """SurfaceScalarCache (src/cache.jl:164-182, _surfacecache :232-261):
    R/E: ScalarData <-> Nodes{Primal}; Rsn/Esn: VectorData <-> Edges{Primal}.
    GridScaling: L.factor = 1/dx^2 (src/cache.jl:323-324)."""

    def __init__(self, grid, x, y, nx, ny, ds, lgf, ddf="yang3", scaling=GRID_SCALING,
                 c0=None, workers=1):
        self.grid = grid
        self.x, self.y = np.asarray(x, float), np.asarray(y, float)
        self.nx, self.ny, self.ds = np.asarray(nx, float), np.asarray(ny, float), np.asarray(ds, float)
        self.N = self.x.shape[0]
        self.scaling, self.ddf = scaling, ddf
        self.tabs = {k: build_table(grid, self.x, self.y, self.ds, k, ddf, scaling)
                     for k in (PRIMAL, XEDGE, YEDGE, DUAL)}
        self.factor = 1.0 / grid.dx ** 2 if scaling == GRID_SCALING else 1.0
        # plan_laplacian(g::PhysicalGrid): DX believed to be cellsize(g) (A.5)
        self.c0 = lgf_c0(grid.dx) if c0 is None else c0
        self.lgf = np.asarray(lgf)[:grid.NX, :grid.NY]
        self.conv = ConvPlan(self.lgf, workers=workers)

    # -- scaling helpers (src/grid_operators.jl:8-9)
    def _scale_derivative(self, w):
        return w / self.grid.dx if self.scaling == GRID_SCALING else w

    # -- basic surface operators (src/surface_operators.jl:13-88)
    def regularize(self, f):
        return regularize(self.tabs[PRIMAL], f)

    def interpolate(self, s):
        return interpolate(self.tabs[PRIMAL], s)

    def regularize_edges(self, fu, fv):
        return regularize(self.tabs[XEDGE], fu), regularize(self.tabs[YEDGE], fv)

    def interpolate_edges(self, u, v):
        return interpolate(self.tabs[XEDGE], u), interpolate(self.tabs[YEDGE], v)

    def regularize_normal(self, f):
        """q = Rf (n o f) (src/surface_operators.jl:100-103)."""
        return self.regularize_edges(self.nx * f, self.ny * f)

    def normal_interpolate(self, u, v):
        """vn = n . Ef q (src/surface_operators.jl:230-233); pointwise_dot =
        n.u*v.u + n.v*v.v."""
        su, sv = self.interpolate_edges(u, v)
        return self.nx * su + self.ny * sv

    def regularize_normal_cross(self, f):
        """q = Rf (n x f e_z): V.u = n.v f, V.v = -n.u f
        (src/cartesian_extensions.jl:19-23)."""
        return self.regularize_edges(self.ny * f, -self.nx * f)

    def normal_cross_interpolate(self, u, v):
        """wn = e_z . (n x Ef q) = n.u*s.v - n.v*s.u."""
        su, sv = self.interpolate_edges(u, v)
        return self.nx * sv - self.ny * su

    # -- grid operators with scaling (src/grid_operators.jl:25-134)
    def divergence(self, u, v):
        return self._scale_derivative(divergence_e2n(self.grid, u, v))

    def grad(self, p):
        u, v = grad_n2e(self.grid, p)
        return self._scale_derivative(u), self._scale_derivative(v)

    def curl_n2e(self, s):
        u, v = curl_n2e(self.grid, s)
        return self._scale_derivative(u), self._scale_derivative(v)

    def curl_e2n(self, u, v):
        return self._scale_derivative(curl_e2n(self.grid, u, v))

    def laplacian(self, w, kind):
        return laplacian(self.grid, w, kind, self.factor)

    def inverse_laplacian(self, w):
        return inverse_laplacian(self.conv, w, self.c0, self.factor)

    # -- composite surface-grid operators (src/surface_operators.jl:357-725)
    def surface_divergence(self, f):
        """theta = D Rf (n o f) / dx (src/surface_operators.jl:530-547)."""
        u, v = self.regularize_normal(f)
        return self._scale_derivative(divergence_e2n(self.grid, u, v))

    def surface_grad(self, phi):
        """vn = n . Ef G phi / dx (src/surface_operators.jl:604-620)."""
        u, v = grad_n2e(self.grid, phi)
        return self._scale_derivative(self.normal_interpolate(u, v))

    def surface_curl_s2n(self, f):
        """w = C^T Rf (n o f) / dx (src/surface_operators.jl:357-375)."""
        u, v = self.regularize_normal(f)
        return self._scale_derivative(curl_e2n(self.grid, u, v))

    def surface_curl_n2s(self, s):
        """vn = n . Ef C s / dx (src/surface_operators.jl:412-429)."""
        u, v = curl_n2e(self.grid, s)
        return self._scale_derivative(self.normal_interpolate(u, v))

    def surface_divergence_cross(self, f):
        """theta = D Rf (n x f e_z)/dx -- the documented operator
        (src/surface_operators.jl:672; the shipped body has a typo, Appendix E)."""
        u, v = self.regularize_normal_cross(f)
        return self._scale_derivative(divergence_e2n(self.grid, u, v))

    def surface_grad_cross(self, phi):
        u, v = grad_n2e(self.grid, phi)
        return self._scale_derivative(self.normal_cross_interpolate(u, v))

    def surface_curl_cross_s2n(self, f):
        u, v = self.regularize_normal_cross(f)
        return self._scale_derivative(curl_e2n(self.grid, u, v))

    def surface_curl_cross_n2s(self, s):
        u, v = curl_n2e(self.grid, s)
        return self._scale_derivative(self.normal_cross_interpolate(u, v))

    # -- masks (src/surface_operators.jl:865-871)
    def mask(self):
        if self.N == 0:
            return np.ones(field_shape(PRIMAL, self.grid.NX, self.grid.NY), order="F")
        m = self.surface_divergence(np.ones(self.N))
        m = self.inverse_laplacian(m)
        return m * -1.0

    # -- Schur builders (src/matrix_operators.jl)
    def mask_product(self, w, kind, complementary=False):
        """mask!(w, cache) / complementary_mask!(w, cache) on a scalar cache
        (src/surface_operators.jl:791-795, 814-818, 880-902): w .*= grid_interpolate(mask)."""
        m = self.mask()
        if complementary:
            m = 1.0 - m                                     # _get_complementary_mask! (:856-859)
        return grid_interpolate(self.grid, m, PRIMAL, kind) * w

    def _probe(self, pre, post, sign, scale, cols=None):
        N = self.N
        cols = range(N) if cols is None else cols
        A = np.zeros((N, len(cols)), order="F")
        for jc, col in enumerate(cols):
            e = np.zeros(N)
            e[col] = 1.0
            g = pre(e)
            g = self.inverse_laplacian(g)
            A[:, jc] = sign * scale * post(g)
        return A

    def apply_schur(self, which, x, scale=1.0):
        """The operator a Schur builder tabulates, applied to a vector: sign*scale*post(L^-1 pre(x)) --
        one column-probe composition (src/matrix_operators.jl:9-30 with e_c replaced by x; the builders
        are linear).  `which`: "RTLinvR", "CLinvCT", "GLinvD", "GLinvD_cross"."""
        pre, post = {"RTLinvR": (self.regularize, self.interpolate),
                     "CLinvCT": (self.surface_curl_s2n, self.surface_curl_n2s),
                     "GLinvD": (self.surface_divergence, self.surface_grad),
                     "GLinvD_cross": (self.surface_divergence_cross, self.surface_grad_cross)}[which]
        return -scale * post(self.inverse_laplacian(pre(np.asarray(x, float))))

    def create_RTLinvR(self, scale=1.0, cols=None):
        """src/matrix_operators.jl:9-30: A[:,c] = -scale*E L^-1 R e_c."""
        return self._probe(self.regularize, self.interpolate, -1.0, scale, cols)

    def create_RTLinvR_table(self, scale=1.0, cols=None):
        """create_RTLinvR without the FFT: table form with the LGF (table_schur)."""
        return table_schur(self.tabs[PRIMAL], self.lgf, c0=self.c0, coef=-scale / self.factor, cols=cols)

    def create_CLinvCT(self, scale=1.0, cols=None):
        """src/matrix_operators.jl:40-61."""
        return self._probe(self.surface_curl_s2n, self.surface_curl_n2s, -1.0, scale, cols)

    def create_GLinvD(self, scale=1.0, cols=None):
        """src/matrix_operators.jl:135-155."""
        return self._probe(self.surface_divergence, self.surface_grad, -1.0, scale, cols)

    def create_GLinvD_cross(self, scale=1.0, cols=None):
        """src/matrix_operators.jl:195-215 (documented operator, Appendix E)."""
        return self._probe(self.surface_divergence_cross, self.surface_grad_cross, -1.0, scale, cols)

    def create_nRTRn(self, scale=1.0):
        """src/matrix_operators.jl:225-244: A[:,c] = scale * n.Ef Rf (n o e_c)."""
        N = self.N
        A = np.zeros((N, N), order="F")
        for col in range(N):
            e = np.zeros(N)
            e[col] = 1.0
            u, v = self.regularize_normal(e)
            A[:, col] = scale * self.normal_interpolate(u, v)
        return A

    def create_surface_filter(self):
        """C = Etilde R (src/matrix_operators.jl:254-268).  Filtered
        interpolation (filter=true upstream): Etilde[k,cell] = E[k,cell] /
        sum_m R[cell,m], zero denominators -> 1, so that C preserves constants
        (medium confidence on upstream's exact form: parity unpinned)."""
        tab = self.tabs[PRIMAL]
        R = R_matrix(tab)
        rowsum = np.asarray(R.sum(axis=1)).ravel()
        rowsum[rowsum == 0.0] = 1.0
        Ef = E_matrix(tab) @ sp.diags(1.0 / rowsum)
        return np.asarray((Ef @ R).todense(), order="F")


def _c_lib():
    """oracle/direct_conv.c built by oracle/Makefile (test infrastructure)."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "libilm_oracle_c.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", here])
    return ctypes.CDLL(so)


def table_schur(tab, K, c0=0.0, coef=1.0, cols=None):
    """A[k,c] = coef * sum_p sum_q wE[k,p] (K(p-q) - c0) wR[c,q]: the regularize -> convolve -> interpolate
    probe of src/matrix_operators.jl:9-30 with the unit-vector loop written out as a table look-up
    (long-double accumulation in oracle/direct_conv.c).  The FFT never enters, which
    makes it the independent full-size check of create_RTLinvR and of the IF-HERK stage complements.
    `K`: kernel table (G for L^-1, the plan_intfact table for exp(L a)); entries beyond it count as zero."""
    import ctypes
    N, W = tab.N, tab.W
    lo, hi = (0, N) if cols is None else cols
    a = np.arange(W)
    ii = np.broadcast_to(tab.i0[:, None, None] + a[None, :, None], (N, W, W))
    jj = np.broadcast_to(tab.j0[:, None, None] + a[None, None, :], (N, W, W))
    pi = np.ascontiguousarray(np.where(tab.valid, ii, -1).reshape(N, W * W), dtype=np.int64)
    pj = np.ascontiguousarray(np.where(tab.valid, jj, -1).reshape(N, W * W), dtype=np.int64)
    wE = np.ascontiguousarray(tab.wE.reshape(N, W * W))
    wR = np.ascontiguousarray(tab.wR.reshape(N, W * W))
    K = np.asfortranarray(K, dtype=np.float64)
    A = np.zeros((N, hi - lo), order="F")
    vp = ctypes.c_void_p
    _c_lib().ilm_oracle_table_schur(vp(K.ctypes.data), K.shape[0], min(K.shape), N, W * W, vp(pi.ctypes.data),
                                    vp(pj.ctypes.data), vp(wE.ctypes.data), vp(wR.ctypes.data), lo, hi,
                                    ctypes.c_double(c0), ctypes.c_double(coef), vp(A.ctypes.data))
    return A


def refined_solve(matvec, b, lu_approx, tol=1e-14, maxit=8):
    """Solve A x = b where A is only available as the ORACLE's operator `matvec` (one probe-like
    composition pre -> L^-1 -> post per application), by iterative refinement preconditioned with the
    LU factors of an approximation of A (any matrix close to A: it steers convergence only, the fixed
    point satisfies the oracle's own equations).  Returns (x, relative residual)."""
    x = scipy.linalg.lu_solve(lu_approx, b)
    res = np.inf
    for _ in range(maxit):
        r = b - matvec(x)
        res = np.abs(r).max() / max(np.abs(b).max(), 1e-300)
        if res < tol:
            break
        x = x + scipy.linalg.lu_solve(lu_approx, r)
    return x, res


class VectorCache(ScalarCache):
    """SurfaceVectorCache (src/cache.jl:184-202): R/E: VectorData <-> Edges{Primal};
    Rsn/Esn: TensorData <-> EdgeGradient{Primal,Dual}; Rcurl/Ecurl: ScalarData <->
    Nodes{Dual}; Rdiv/Ediv: ScalarData <-> Nodes{Primal}.  VectorData is (u, v);
    TensorData / EdgeGradient are (dudx, dudy, dvdx, dvdy) with dudx, dvdy on primal
    nodes and dudy, dvdx on dual nodes (src/cache.jl:682-690).

    Tensor component convention (PARITY UNPINNED): component (i, j)
    = (velocity component i, derivative direction j); pointwise_tensorproduct!(T, n, v)
    gives T.d(v_i)/d(x_j) = v_i n_j and pointwise_dot!(t, n, T) gives t_i = sum_j n_j T_ij,
    the pair that makes regularize_normal!/normal_interpolate! adjoint and D R_t (n o v)
    a double layer."""

    # -- pointwise tensor algebra
    def tensorproduct(self, vu, vv):
        n1, n2 = self.nx, self.ny
        return n1 * vu, n2 * vu, n1 * vv, n2 * vv

    def tensordot(self, T):
        n1, n2 = self.nx, self.ny
        return n1 * T[0] + n2 * T[1], n1 * T[2] + n2 * T[3]

    # -- TensorData <-> EdgeGradient tables
    def regularize_tensor(self, T):
        return (regularize(self.tabs[PRIMAL], T[0]), regularize(self.tabs[DUAL], T[1]),
                regularize(self.tabs[DUAL], T[2]), regularize(self.tabs[PRIMAL], T[3]))

    def interpolate_tensor(self, A):
        return (interpolate(self.tabs[PRIMAL], A[0]), interpolate(self.tabs[DUAL], A[1]),
                interpolate(self.tabs[DUAL], A[2]), interpolate(self.tabs[PRIMAL], A[3]))

    # -- src/surface_operators.jl:113-138
    def regularize_normal_v(self, vu, vv):
        return self.regularize_tensor(self.tensorproduct(vu, vv))

    def symm_tensor(self, vu, vv):
        t = self.tensorproduct(vu, vv)
        dot = self.nx * vu + self.ny * vv
        return (t[0] + t[0]) - dot, t[1] + t[2], t[2] + t[1], (t[3] + t[3]) - dot

    def regularize_normal_symm(self, vu, vv):
        return self.regularize_tensor(self.symm_tensor(vu, vv))

    # -- src/surface_operators.jl:242-267
    def normal_interpolate_v(self, A):
        return self.tensordot(self.interpolate_tensor(A))

    def normal_interpolate_symm(self, A):
        tr = A[0] + A[3]
        Gs = ((A[0] + A[0]) - tr, A[2] + A[1], A[1] + A[2], (A[3] + A[3]) - tr)
        return self.tensordot(self.interpolate_tensor(Gs))

    # -- src/surface_operators.jl:172-215, 303-343
    def regularize_normal_cross_v(self, vu, vv):
        """w = Rcurl (n x v), pointwise_cross!(s,n,v) = n.u v.v - n.v v.u."""
        return regularize(self.tabs[DUAL], self.nx * vv - self.ny * vu)

    def regularize_normal_dot_v(self, vu, vv):
        return regularize(self.tabs[PRIMAL], self.nx * vu + self.ny * vv)

    def regularize_normal_dot_t(self, T):
        tu, tv = self.tensordot(T)
        return self.regularize_edges(tu, tv)

    def normal_cross_interpolate_v(self, s):
        f = interpolate(self.tabs[DUAL], s)
        return self.ny * f, -self.nx * f

    def normal_dot_interpolate_v(self, phi):
        f = interpolate(self.tabs[PRIMAL], phi)
        return self.nx * f, self.ny * f

    def normal_dot_interpolate_t(self, u, v):
        su, sv = self.interpolate_edges(u, v)
        return self.tensorproduct(su, sv)

    # -- composites (src/surface_operators.jl:388-398, 443-452, 559-591, 632-664)
    def surface_divergence_v(self, vu, vv, symm=False):
        T = self.regularize_normal_symm(vu, vv) if symm else self.regularize_normal_v(vu, vv)
        u, v = divergence_t2e(self.grid, *T)
        return self._scale_derivative(u), self._scale_derivative(v)

    def surface_grad_v(self, u, v, symm=False):
        A = grad_e2t(self.grid, u, v)
        tu, tv = self.normal_interpolate_symm(A) if symm else self.normal_interpolate_v(A)
        return self._scale_derivative(tu), self._scale_derivative(tv)

    def surface_curl_v2n(self, vu, vv):
        u, v = self.regularize_edges(vu, vv)
        return self._scale_derivative(curl_e2n(self.grid, u, v))

    def surface_curl_n2v(self, s):
        u, v = curl_n2e(self.grid, s)
        fu, fv = self.interpolate_edges(u, v)
        return self._scale_derivative(fu), self._scale_derivative(fv)

    def inverse_laplacian_edges(self, u, v):
        return self.inverse_laplacian(u), self.inverse_laplacian(v)

    def mask_edges(self):
        """_get_mask! with Edges grid data (src/surface_operators.jl:865-871)."""
        one = np.ones(self.N)
        u, v = self.surface_divergence_v(one, one)
        u, v = self.inverse_laplacian_edges(u, v)
        return u * -1.0, v * -1.0

    # -- Schur builders on VectorData (2N x 2N), src/matrix_operators.jl
    def mask_product_v(self, w, kind, complementary=False):
        """mask!/complementary_mask! on a vector cache (:796-800, 819-823, 904-923).  `w` is a tuple
        (u, v) for kind 'edges', 4 components for 'edgegrad', one array for PRIMAL / DUAL."""
        mu, mv = self.mask_edges()
        if complementary:
            mu, mv = 1.0 - mu, 1.0 - mv
        if kind == "edges":
            return mu * w[0], mv * w[1]
        if kind == "edgegrad":                               # dudx, dvdy primal; dudy, dvdx dual; all from mask.u
            kinds = (PRIMAL, DUAL, DUAL, PRIMAL)
            return tuple(grid_interpolate(self.grid, mu, XEDGE, k) * c for k, c in zip(kinds, w))
        return grid_interpolate(self.grid, mu, XEDGE, kind) * w

    def _probe_v(self, pre, post, sign, scale, nsolve=1, cols=None):
        N = self.N
        cols = range(2 * N) if cols is None else cols
        A = np.zeros((2 * N, len(cols)), order="F")
        for jc, col in enumerate(cols):
            e = np.zeros(2 * N)
            e[col] = 1.0
            g = pre(e[:N], e[N:])
            for _ in range(nsolve):
                g = tuple(self.inverse_laplacian(c) for c in g) if isinstance(g, tuple) else self.inverse_laplacian(g)
            out = post(*g) if isinstance(g, tuple) else post(g)
            A[:, jc] = sign * scale * np.concatenate(out)
        return A

    def apply_CL2invCT(self, x, scale=1.0):
        """create_CL2invCT's operator applied to a VectorData x = [u; v] (src/matrix_operators.jl:102-125)."""
        N = self.N
        g = self.surface_curl_v2n(x[:N], x[N:])
        g = self.inverse_laplacian(self.inverse_laplacian(g))
        return -scale * np.concatenate(self.surface_curl_n2v(g))

    def create_RTLinvR_v(self, scale=1.0, cols=None):
        return self._probe_v(self.regularize_edges, self.interpolate_edges, -1.0, scale, cols=cols)

    def create_CLinvCT_v(self, scale=1.0, cols=None):
        """create_CLinvCT on a vector cache (src/matrix_operators.jl:40-61)."""
        return self._probe_v(self.surface_curl_v2n, self.surface_curl_n2v, -1.0, scale, cols=cols)

    def create_CL2invCT(self, scale=1.0, cols=None):
        """src/matrix_operators.jl:102-125: two inverse Laplacians."""
        return self._probe_v(self.surface_curl_v2n, self.surface_curl_n2v, -1.0, scale, nsolve=2, cols=cols)

    def create_GLinvD_v(self, scale=1.0, symm=False, cols=None):
        """create_GLinvD / create_GLinvD_symm on a vector cache (:135-185)."""
        return self._probe_v(lambda a, b: self.surface_divergence_v(a, b, symm),
                             lambda u, v: self.surface_grad_v(u, v, symm), -1.0, scale, cols=cols)

    def create_nRTRn_v(self, scale=1.0):
        N = self.N
        A = np.zeros((2 * N, 2 * N), order="F")
        for col in range(2 * N):
            e = np.zeros(2 * N)
            e[col] = 1.0
            A[:, col] = scale * np.concatenate(self.normal_interpolate_v(self.regularize_normal_v(e[:N], e[N:])))
        return A

    # create_CLinvCT_scalar (:71-92) on a vector cache is the scalar-cache create_CLinvCT
    create_CLinvCT_scalar = ScalarCache.create_CLinvCT


def dirichlet_solve(cache, fplus, fminus=None, S=None, solve=None):
    """Block-LU Dirichlet Poisson (test/literate/dirichlet.jl:71-107).
    Returns (f, s, S).  `solve`: optional b -> S^-1 b (e.g. refined_solve on apply_schur) used instead of
    building and factoring S -- the full-size checks cannot afford N oracle probes."""
    fminus = np.zeros_like(fplus) if fminus is None else fminus
    d = fplus - fminus
    fb = 0.5 * (fplus + fminus)
    fstar = cache.surface_divergence(d)
    fstar = cache.inverse_laplacian(fstar)
    if S is None and solve is None:
        S = cache.create_RTLinvR()
    s = cache.interpolate(fstar)
    s = fb - s
    s = -solve(s) if solve is not None else -scipy.linalg.lu_solve(scipy.linalg.lu_factor(S), s)
    f = cache.regularize(s)
    f = cache.inverse_laplacian(f)
    f = f + fstar
    return f, s, S


def neumann_solve(cache, vnplus, vnminus=None, S=None, solve=None):
    """Block-LU Neumann Poisson with streamfunction (test/literate/neumann.jl:101-142).
    Returns (f, df, s, ds, S)."""
    vnminus = np.zeros_like(vnplus) if vnminus is None else vnminus
    dvn = vnplus - vnminus
    vn = 0.5 * (vnplus + vnminus)
    if solve is None:
        if S is None:
            S = cache.create_CLinvCT()
        lu = scipy.linalg.lu_factor(S)
        solve = lambda b: scipy.linalg.lu_solve(lu, b)      # noqa: E731
    fstar = cache.inverse_laplacian(cache.regularize(dvn))
    df = cache.surface_grad(fstar)
    df = vn - df
    df = -solve(df)
    f = cache.inverse_laplacian(cache.surface_divergence(df)) + fstar
    sstar = cache.surface_curl_s2n(df)
    ds = solve(cache.surface_grad_cross(fstar))
    s = cache.surface_curl_cross_s2n(ds)
    s = (s - sstar) * -1.0
    s = cache.inverse_laplacian(s)
    return f, df, s, ds, S


# --------------------------------------------------------------------------
# inner products of the reference used by the pinning tests
# --------------------------------------------------------------------------
def dot_grid(grid, a, b, kind):
    """Grid inner product weighted by dx^2 (src/tools.jl:101-104).  Upstream's
    dot is a trapezoid rule over the physical domain: along a primal direction
    the first/last entries weigh 1/2 (Appendix B `integrate(ones_grid)` =
    404^2 dx^2 on 405^2 primal nodes), along a dual direction the ghost entries
    weigh 0 (test/tools.jl:37: norm(1,g)^2 = domain area on dual nodes)."""
    sx, sy = field_shift(kind)
    mx, my = a.shape

    def w1(m, dual):
        w = np.ones(m)
        w[0] = w[-1] = 0.0 if dual else 0.5
        return w
    w = w1(mx, sx != 0.0)[:, None] * w1(my, sy != 0.0)[None, :]
    return float(np.sum(w * a * b) * grid.dx ** 2)


def dot_surface(a, b, ds):
    return float(np.sum(a * b * ds))


# --------------------------------------------------------------------------
# IF-HERK step of the constrained heat equation (config C3)
# --------------------------------------------------------------------------
def heat_ifherk_step(cache, T, t, dt, kappa, tab_a, tab_c, tables, Tplus, Tminus, reuse=None, schur="probe"):
    """One step of the half-explicit Runge-Kutta scheme with integrating factor that the reference
    reaches through ConstrainedSystems.jl's LiskaIFHERK (src/timemarching.jl:86-107,254-260;
    problem functions of test/literate/heatconduction.jl:87-129):
        dT/dt = kappa L T + D_s(-kappa [T]) - R sigma,   E T = (T+ + T-)/2.
    `tables[a]` is the plan_intfact table for the argument a = kappa/dx^2 (c_i - c_{i-1}) dt (a = 0:
    identity).  The integrator itself is un-vendored; the recursion is PINNED by the reference's executed notebook
    examples/heatconduction.ipynb (cells 58, 62): T(-0.9, 0) after 51 and 54 steps is reproduced to 2e-14
    (tests/test_golden.py::test_notebook_heatconduction_time_marching).
    `reuse`: a dict that keeps the convolution plans and the LU factors of the stage complements between steps
    (static body); None rebuilds them every step (moving body).
    `schur`: "probe" builds S_i column by column as the reference does; "table" uses the FFT-free table form
    (table_schur), the same matrix up to rounding, affordable at BASELINE sizes (2048^2, N = 2295)."""
    g = cache.grid
    N = cache.N
    tp = cache.tabs[PRIMAL]

    def sval(f, tt):
        return np.asarray(f(cache.x, cache.y, tt), dtype=float) if callable(f) else np.full(N, float(f))

    def H(wf, a):
        if a == 0.0:
            return wf
        if reuse is None:
            return ConvPlan(tables[a][:g.NX, :g.NY], workers=getattr(cache.conv, "workers", 1)).apply(wf)
        if ("plan", a) not in reuse:
            reuse[("plan", a)] = ConvPlan(tables[a][:g.NX, :g.NY], workers=getattr(cache.conv, "workers", 1))
        return reuse[("plan", a)].apply(wf)

    def rhs(tt):
        return cache.surface_divergence(-kappa * (sval(Tplus, tt) - sval(Tminus, tt)))

    Emat, Rmat = E_matrix(tp), R_matrix(tp)
    q = T.copy()
    w = []
    c_prev = 0.0
    U = None
    sig = None
    for i, c in enumerate(tab_c):
        a = kappa / g.dx ** 2 * (c - c_prev) * dt
        w = [H(wj, a) for wj in w]
        q = H(q, a)
        w.append(H(rhs(t + c_prev * dt), a))
        U = q.copy()
        for j in range(i + 1):
            if tab_a[i][j] != 0.0:
                U = U + (dt * tab_a[i][j]) * w[j]
        # S_i = -E H_i R, column by column like create_RTLinvR (src/matrix_operators.jl:9-30)
        if reuse is None or ("lu", a) not in reuse:
            if schur == "table":
                S = table_schur(tp, tables[a][:g.NX, :g.NY] if a != 0.0 else np.ones((1, 1)), coef=-1.0)
            else:
                S = np.zeros((N, N))
                for col in range(N):
                    e = np.zeros(N)
                    e[col] = 1.0
                    S[:, col] = -interpolate(tp, H(regularize(tp, e), a))
            if reuse is not None:
                reuse[("lu", a)] = scipy.linalg.lu_factor(S)
        b = 0.5 * (sval(Tplus, t + c * dt) + sval(Tminus, t + c * dt))
        if reuse is None:
            sig = np.linalg.solve(S, b - interpolate(tp, U))
        else:
            sig = scipy.linalg.lu_solve(reuse[("lu", a)], b - interpolate(tp, U))
        corr = H(regularize(tp, sig), a)
        U = U - corr
        w[i] = w[i] - corr * (1.0 / (dt * tab_a[i][i]))
        c_prev = c
    del Emat, Rmat
    return U, sig


def stokes_solve(cache, vplus, vminus=None, solve_S=None, solve_Ss=None):
    """`solve(prob::StokesFlowProblem, sys)` of test/literate/stokes.jl:98-166 on a VectorCache (without the
    final `C^2 * sigma` traction filter).  vplus / vminus: (2N,) [u; v].  Returns (vu, vv, s, sigma)."""
    N = cache.N
    vplus = np.asarray(vplus, float)
    vminus = np.zeros(2 * N) if vminus is None else np.asarray(vminus, float)
    dv = vplus - vminus
    vu, vv = cache.surface_divergence_v(dv[:N], dv[N:], symm=True)
    sstar = -1.0 * cache.curl_e2n(vu, vv)
    sstar = cache.inverse_laplacian(cache.inverse_laplacian(sstar))
    dvn = cache.nx * dv[:N] + cache.ny * dv[N:]
    phi = cache.inverse_laplacian(cache.regularize(dvn))
    pu, pv = cache.grad(phi)
    eu, ev = cache.interpolate_edges(pu, pv)
    vprime = 0.5 * (vplus + vminus) - np.concatenate([eu, ev])
    cu, cv = cache.curl_n2e(sstar)
    eu, ev = cache.interpolate_edges(cu, cv)
    vprime = vprime - np.concatenate([eu, ev])
    sigma = solve_S(vprime) if solve_S is not None else np.linalg.solve(cache.create_CL2invCT(), vprime)
    s = -1.0 * cache.surface_curl_v2n(sigma[:N], sigma[N:])
    s = cache.inverse_laplacian(cache.inverse_laplacian(s)) + sstar
    cu, cv = cache.curl_n2e(s)
    vu, vv = cu + pu, cv + pv
    ds = cache.surface_grad_cross(phi)
    ds = solve_Ss(ds) if solve_Ss is not None else np.linalg.solve(cache.create_CLinvCT(), ds)
    s = s + cache.inverse_laplacian(-1.0 * cache.surface_curl_cross_s2n(ds))
    return vu, vv, s, sigma


# --------------------------------------------------------------------------
# convective terms (src/grid_operators.jl:258-434), composed exactly as the reference composes them
# --------------------------------------------------------------------------
def convective_derivative_scalar(grid, u, v, p, div=1.0):
    """_unscaled_convective_derivative!(udp, u, p) (:318-327) then _scale_derivative!:
    grad!(vt1, p); product!(vt3, u, vt1); grid_interpolate!(udp, vt3) (x- and y-edge parts summed)."""
    gu, gv = grad_n2e(grid, p)
    out = grid_interpolate(grid, u * gu, XEDGE, PRIMAL) + grid_interpolate(grid, v * gv, YEDGE, PRIMAL)
    return out / div


def w_cross_v(grid, w, u, v):
    """_unscaled_w_cross_v! (:419-434)."""
    t1 = -1.0 * grid_interpolate(grid, v, YEDGE, DUAL)
    ou = grid_interpolate(grid, t1 * w, DUAL, XEDGE)
    t1 = grid_interpolate(grid, u, XEDGE, DUAL)
    ov = grid_interpolate(grid, t1 * w, DUAL, YEDGE)
    return ou, ov


def convective_derivative_vector(grid, cu, cv, u, v, div=1.0):
    """_unscaled_convective_derivative!(vdu, v, u) (:361-375; (:343-359) when u is v): grid_interpolate!
    of the advecting velocity to the EdgeGradient positions, transpose!, grad!(u), product!,
    grid_interpolate! back to the edges (the two contributions of each component summed)."""
    a0 = grid_interpolate(grid, cu, XEDGE, PRIMAL)      # vt1.dudx
    a1 = grid_interpolate(grid, cu, XEDGE, DUAL)        # vt1.dudy
    a2 = grid_interpolate(grid, cv, YEDGE, DUAL)        # vt1.dvdx
    a3 = grid_interpolate(grid, cv, YEDGE, PRIMAL)      # vt1.dvdy
    dudx, dudy, dvdx, dvdy = grad_e2t(grid, u, v)
    p0, p1, p2, p3 = a0 * dudx, a2 * dudy, a1 * dvdx, a3 * dvdy      # transpose! swaps the off-diagonal slots
    ou = grid_interpolate(grid, p0, PRIMAL, XEDGE) + grid_interpolate(grid, p1, DUAL, XEDGE)
    ov = grid_interpolate(grid, p2, DUAL, YEDGE) + grid_interpolate(grid, p3, PRIMAL, YEDGE)
    return ou / div, ov / div


def convective_derivative_dual(grid, u, v, w, div=1.0):
    """_unscaled_convective_derivative!(udw::Nodes{Dual}, u, w) (:329-341) with Edges{Dual} temporaries (x-component
    at the primal y-edge positions, y-component at the primal x-edge positions): grad!(vt1, w);
    grid_interpolate!(vt2, u) (4-point means); product!; grid_interpolate!(udw, vt3).  Parity unpinned like
    grid_interpolate itself."""
    gx = w[1:, :] - w[:-1, :]                                  # (NX-1) x NY, at (i, j-1/2)
    gy = w[:, 1:] - w[:, :-1]                                  # NX x (NY-1), at (i-1/2, j)
    cu = grid_interpolate(grid, u, XEDGE, YEDGE)
    cv = grid_interpolate(grid, v, YEDGE, XEDGE)
    out = grid_interpolate(grid, cu * gx, YEDGE, DUAL) + grid_interpolate(grid, cv * gy, XEDGE, DUAL)
    return out / div



# --------------------------------------------------------------------------
# forcing regions (src/forcing.jl:456-515)
# --------------------------------------------------------------------------
def point_collection_table(grid, x, y, kind, ddf="yang3"):
    """PointCollectionCache's regop (src/cache.jl:269-292, 312-313): Regularize(X, dx, I0 = origin, ddftype) with the
    default weight 1 and no symmetry, i.e. wR = ddf ddf / dx^2."""
    x = np.asarray(x, float)
    return build_table(grid, x, np.asarray(y, float), np.ones_like(x), kind, ddf, GRID_SCALING)


def forcing_area(dy, strength, mask=None):
    """_apply_forcing! on an AreaRegionCache (src/forcing.jl:456-465): dy .+= str .* mask; the whole-domain region
    has mask = ones (_get_mask! on a cache without points, src/surface_operators.jl:873-876)."""
    prod = strength if mask is None else strength * mask
    return dy + prod


def forcing_line(dy, tab, strength):
    """_apply_forcing! on a LineRegionCache / PointRegionCache (src/forcing.jl:467-494): fill!(gdata_cache, 0);
    regularize!(gdata_cache, str); dy .+= gdata_cache."""
    return dy + regularize(tab, strength)


# --------------------------------------------------------------------------
# Helmholtz decomposition on a VectorCache (src/helmholtz.jl)
# --------------------------------------------------------------------------
def helmholtz_jump(vc, op, sign, dvu, dvv, field):
    """op 'cross' on Nodes{Dual}: field + sign * R (n x dv); op 'dot' on Nodes{Primal}: field + sign * R (n . dv).
    sign -1: masked_curlv_from_curlv_masked! / masked_divv_from_divv_masked! (src/helmholtz.jl:149-160, 240-252: fill,
    regularize, .*= -1, .+= field); sign +1: curlv_masked_from_masked_curlv! / divv_masked_from_masked_divv!
    (:171-181, 263-274).  A cache without points passes the field through (VectorData{0} methods)."""
    if vc.N == 0:
        return np.array(field, copy=True)
    r = vc.regularize_normal_cross_v(dvu, dvv) if op == "cross" else vc.regularize_normal_dot_v(dvu, dvv)
    if sign < 0:
        r = r * -1.0
    return r + field


def helmholtz_potentials(vc, curlv, divv, dvu, dvv):
    """vectorpotential_from_masked_curlv! (src/helmholtz.jl:84-96, 114-124): stemp = R n x dv; stemp .+= curlv;
    psi = -(L^-1 stemp).  scalarpotential_from_masked_divv! (:186-201, 216-219): phi = L^-1 (R n . dv + divv)."""
    psi = vc.inverse_laplacian(helmholtz_jump(vc, "cross", +1, dvu, dvv, curlv)) * -1.0
    phi = vc.inverse_laplacian(helmholtz_jump(vc, "dot", +1, dvu, dvv, divv))
    return psi, phi


def vecfield_from_potentials(vc, psi, phi, vp=None):
    """vecfield_from_vectorpotential! (curl!, :130-134), vecfield_from_scalarpotential! (grad!, :228-232) and the sum
    of vecfield_helmholtz! (:300-304): v .= vpsi .+ vphi; v .+= vp."""
    cu, cv = vc.curl_n2e(psi)
    gu, gv = vc.grad(phi)
    u, v = cu + gu, cv + gv
    if vp is not None:
        u, v = u + vp[0], v + vp[1]
    return u, v


def vecfield_helmholtz(vc, curlv, divv, dvu, dvv, vp=None):
    """vecfield_helmholtz! (src/helmholtz.jl:285-307)."""
    psi, phi = helmholtz_potentials(vc, curlv, divv, dvu, dvv)
    return vecfield_from_potentials(vc, psi, phi, vp)


def heat_unbounded_step(grid, T, t, dt, kappa, tab_a, tab_c, tables, rhs):
    """One integrating-factor Runge-Kutta step of the unconstrained heat equation dT/dt = kappa L T + rhs(T, t)
    (test/literate/heatconduction-unbounded.jl: ode_rhs = heatconduction_rhs!, lin_op = heat_L, no constraint), the
    recursion of heat_ifherk_step without the multiplier.  PARITY UNPINNED (un-vendored integrator)."""
    def H(wf, a):
        return wf if a == 0.0 else ConvPlan(tables[a][:grid.NX, :grid.NY]).apply(wf)

    q = T.copy()
    U = T
    w = []
    c_prev = 0.0
    for i, c in enumerate(tab_c):
        a = kappa / grid.dx ** 2 * (c - c_prev) * dt
        r = rhs(U, t + c_prev * dt)
        w = [H(wj, a) for wj in w]
        q = H(q, a)
        w.append(H(r, a))
        U = q.copy()
        for j in range(i + 1):
            if tab_a[i][j] != 0.0:
                U = U + (dt * tab_a[i][j]) * w[j]
        c_prev = c
    return U

"""Lattice Green's function tables (host-side input data for the plan).

The C ABI takes the LGF table as *data* (SURVEY.md fact 7): the Julia shim
passes CartesianGrids' `LGF_TABLE`; standalone runs generate one here and hand
the same numbers to the GPU plan and to the test oracle.

Two documented rules:

``rule="accurate"`` (default)
    G(i,j) = (1/2pi) int_0^pi [1 - exp(-i a(k)) cos(j k)] / sinh a(k) dk,
    cosh a(k) = 2 - cos k, for i >= j >= 0 (symmetric otherwise), evaluated with
    composite Gauss-Legendre panels for max(i,j) <= R_QUAD, and the three-term
    far-field expansion
        G ~ (ln r + gamma + 0.5 ln 8)/(2 pi) - cos(4t)/(24 pi r^2)
            - (18 cos(4t) + 25 cos(8t))/(480 pi r^4)
    beyond.  Sign convention L G = +delta, G(0,0) = 0.

``rule="gl100"``
    the reference-compatible rule: the t-integral of A.4 with one fixed
    100-node Gauss-Legendre rule on (0,1) (what upstream tabulates for indices
    < 1600; its quadrature error grows to 1e-9 beyond index ~400).
"""
from __future__ import annotations

import os

import numpy as np

EULER_GAMMA = 0.57721566490153286
R_QUAD = 192


def _asymptotic(i, j):
    i = np.asarray(i, dtype=np.float64)
    j = np.asarray(j, dtype=np.float64)
    r2 = i * i + j * j
    th = np.arctan2(j, i)
    c4, c8 = np.cos(4 * th), np.cos(8 * th)
    return ((0.5 * np.log(r2) + EULER_GAMMA + 0.5 * np.log(8.0)) / (2 * np.pi)
            - c4 / (24 * np.pi * r2)
            - (18 * c4 + 25 * c8) / (480 * np.pi * r2 * r2))


def _quad_rows(n):
    """G(i,j) for 0 <= j <= i < n by composite Gauss-Legendre."""
    npan = max(16, 4 * n)
    xg, wg = np.polynomial.legendre.leggauss(24)
    # uniform panels resolve cos(jk); geometric panels towards k=0 resolve the
    # 1/k boundary layer of width ~1/i
    edges = np.unique(np.concatenate([np.linspace(0.0, np.pi, npan + 1),
                                      np.pi * 2.0 ** -np.arange(1, 30)]))
    h = np.diff(edges) / 2
    k = (edges[:-1, None] + h[:, None] * (xg[None, :] + 1)).ravel()
    w = (h[:, None] * wg[None, :]).ravel()
    a = 2.0 * np.arcsinh(np.sin(0.5 * k))   # cosh a = 2 - cos k, stable near k=0
    sh = np.sinh(a)
    G = np.zeros((n, n))
    jj = np.arange(n)
    cosjk = np.cos(jj[:, None] * k[None, :])          # (n, K)
    for i in range(n):
        e = np.exp(-i * a)
        integrand = (1.0 - e[None, :] * cosjk[: i + 1]) / sh[None, :]
        G[i, : i + 1] = integrand @ w / (2 * np.pi)
    G = np.tril(G) + np.tril(G, -1).T
    G[0, 0] = 0.0
    return G


def _gl100(n):
    xg, wg = np.polynomial.legendre.leggauss(100)
    t = 0.5 * (xg + 1.0)
    w = 0.5 * wg
    si, sm = np.sqrt(1j), np.sqrt(-1j)
    A = (t - si) / (t + si)
    B = (t + sm) / (t - sm)
    G = np.zeros((n, n))
    # powers built incrementally: row i needs A^(j+i) B^(j-i), j<=i
    for i in range(n):
        j = np.arange(i + 1)
        val = 1.0 - A[None, :] ** (j[:, None] + i) * B[None, :] ** (j[:, None] - i)
        G[i, : i + 1] = (np.real(val) / t[None, :]) @ w / (2 * np.pi)
    G = np.tril(G) + np.tril(G, -1).T
    G[0, 0] = 0.0
    return G


def lgf_table(n, rule="accurate", cache_dir=None):
    """(n, n) table G[i, j], 0 <= i, j < n."""
    n = int(n)
    if cache_dir is not None:
        fn = os.path.join(cache_dir, f"lgf_{rule}_{n}.npy")
        if os.path.exists(fn):
            return np.load(fn)
    if rule == "gl100":
        G = _gl100(n)
    elif rule == "accurate":
        nq = min(n, R_QUAD)
        G = np.empty((n, n))
        if n > nq:
            i = np.arange(n)
            with np.errstate(divide="ignore", invalid="ignore"):
                G[:, :] = _asymptotic(i[:, None], i[None, :])
        G[:nq, :nq] = _quad_rows(nq)
    else:
        raise ValueError(f"unknown LGF rule {rule!r}")
    if cache_dir is not None:
        os.makedirs(cache_dir, exist_ok=True)
        np.save(fn, G)
    return G


def lgf_c0(DX=1.0):
    """Far-field constant subtracted by upstream's `L\\w`."""
    return (EULER_GAMMA + 0.5 * np.log(8.0) - np.log(DX)) / (2.0 * np.pi)


def intfact_table(a, n):
    """plan_intfact kernel E_a(i,j) = exp(-4a) I_i(2a) I_j(2a), zeroed beyond the
    first index where E_a(Nmax,0) < eps."""
    from scipy.special import ive
    i = np.arange(n)
    # exp(-2a) I_i(2a) = ive(i, 2a)
    col = ive(i, 2.0 * a)
    E = col[:, None] * col[None, :]
    below = np.nonzero(E[:, 0] < np.finfo(np.float64).eps)[0]
    if below.size:
        nmax = below[0]
        E[nmax + 1:, :] = 0.0
        E[:, nmax + 1:] = 0.0
    return E


def implicit_diffusion_table(a, n):
    """plan_implicit_diffusion kernel: impulse response of (I - a L)^-1 on the infinite lattice
    (L = unscaled 5-point Laplacian), `implicit_operator(L, a)` of src/grid_operators.jl:182-184.
    K(m,n) = (1/pi) int_0^pi cos(m k) rho(k)^|n| / sqrt(A(k)^2 - 4 a^2) dk with A = 1 + 4a - 2a cos k,
    rho = (A - sqrt(A^2 - 4a^2)) / (2a)  (the y-sum done in closed form); the periodic integrand is
    analytic, so the trapezoid rule (an FFT) converges geometrically.  Entries below eps*K(0,0) are
    zeroed like plan_intfact does.  Upstream's table is not recalled: parity unpinned."""
    n = int(n)
    if a == 0:
        K = np.zeros((n, n))
        K[0, 0] = 1.0
        return K
    M = 8 * max(n, 64)
    k = 2.0 * np.pi * np.arange(M) / M
    A = 1.0 + 4.0 * a - 2.0 * a * np.cos(k)
    root = np.sqrt(A * A - 4.0 * a * a)
    rho = (A - root) / (2.0 * a)
    K = np.zeros((n, n))
    p = 1.0 / root
    for j in range(n):
        col = np.fft.ifft(p).real[:n]
        K[:, j] = col
        p = p * rho
        if np.abs(col).max() < np.finfo(np.float64).eps * 1e-3:
            break
    K = 0.5 * (K + K.T)
    K[np.abs(K) < np.finfo(np.float64).eps * K[0, 0]] = 0.0
    return K

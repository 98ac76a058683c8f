"""Synthetic immersed bodies for tests and benchmarks (host-side input synthesis).

Midpoint-rule polygons following the convention recorded in SURVEY.md Appendix B
for RigidBodyTools' shapes: `Circle(R, ds)` has N = floor(2 pi R / ds) points that
lie ON the circle at theta_k = 2 pi k / N, outward normals, and arc weight
ds_k = 2 R tan(pi/N) (the side of the circumscribing polygon), which reproduces
`sum(ds) = 6.283288300933757` for R=1, N=448 (examples/caches.ipynb:1380).
Every function returns (x, y, nx, ny, ds) as float64 arrays.
"""
from __future__ import annotations

import numpy as np


def circle(radius, ds_target, center=(0.0, 0.0)):
    n = int(np.floor(2.0 * np.pi * radius / ds_target))
    th = 2.0 * np.pi * np.arange(n) / n
    nx, ny = np.cos(th), np.sin(th)
    x = center[0] + radius * nx
    y = center[1] + radius * ny
    ds = np.full(n, 2.0 * radius * np.tan(np.pi / n))
    return x, y, nx, ny, ds


def ellipse(a, b, ds_target, center=(0.0, 0.0)):
    """Polygon with vertices equally spaced in arc length (numerically) on the
    ellipse; points are segment midpoints, normals the segment normals."""
    # dense arc-length parametrisation
    m = 200001
    t = np.linspace(0.0, 2.0 * np.pi, m)
    xe, ye = a * np.cos(t), b * np.sin(t)
    seg = np.hypot(np.diff(xe), np.diff(ye))
    s = np.concatenate([[0.0], np.cumsum(seg)])
    n = int(np.floor(s[-1] / ds_target))
    sv = s[-1] * np.arange(n + 1) / n
    tv = np.interp(sv, s, t)
    xv, yv = a * np.cos(tv), b * np.sin(tv)
    return _polygon_midpoints(xv, yv, center)


def rectangle(ha, hb, ds_target, center=(0.0, 0.0)):
    """Rectangle with half-lengths (ha, hb), counter-clockwise."""
    nxs = max(int(np.round(2 * ha / ds_target)), 1)
    nys = max(int(np.round(2 * hb / ds_target)), 1)
    bx = np.linspace(-ha, ha, nxs + 1)
    by = np.linspace(-hb, hb, nys + 1)
    xv = np.concatenate([bx[:-1], np.full(nys, ha), bx[::-1][:-1], np.full(nys, -ha)])
    yv = np.concatenate([np.full(nxs, -hb), by[:-1], np.full(nxs, hb), by[::-1][:-1]])
    xv = np.append(xv, xv[0])
    yv = np.append(yv, yv[0])
    return _polygon_midpoints(xv, yv, center)


def plate(length, ds_target, center=(0.0, 0.0), angle=0.0):
    """Open flat plate: n segments, midpoints, normal = +90 deg from tangent."""
    n = max(int(np.round(length / ds_target)), 1)
    sv = np.linspace(-0.5 * length, 0.5 * length, n + 1)
    c, s = np.cos(angle), np.sin(angle)
    xv, yv = c * sv, s * sv
    xm = 0.5 * (xv[1:] + xv[:-1]) + center[0]
    ym = 0.5 * (yv[1:] + yv[:-1]) + center[1]
    ds = np.hypot(np.diff(xv), np.diff(yv))
    nx = np.full(n, -s)
    ny = np.full(n, c)
    return xm, ym, nx, ny, ds


def _polygon_midpoints(xv, yv, center):
    dx, dy = np.diff(xv), np.diff(yv)
    ds = np.hypot(dx, dy)
    xm = 0.5 * (xv[1:] + xv[:-1]) + center[0]
    ym = 0.5 * (yv[1:] + yv[:-1]) + center[1]
    # counter-clockwise traversal -> outward normal = (dy, -dx)/ds
    return xm, ym, dy / ds, -dx / ds, ds


def concat(*bodies):
    """BodyList surrogate: concatenated arrays plus body_first offsets (nb+1)."""
    arrs = [np.concatenate([b[i] for b in bodies]) for i in range(5)]
    first = np.concatenate([[0], np.cumsum([b[0].shape[0] for b in bodies])]).astype(np.int32)
    return (*arrs, first)


def multibody_c4(dx, radius=0.1):
    """BASELINE config C4 ("8 cylinders + plate, ~4000 surface points" on 4096^2):
    8 circles on a ring of radius 1.2 plus a flat plate of length 1 at the
    centre, ds = 1.4 dx.  radius = 0.1 gives N = 4406 at dx = 4/4094 (SURVEY's
    R = 0.25 would give 9915 points, not ~4000)."""
    ds = 1.4 * dx
    bl = []
    for k in range(8):
        th = 2.0 * np.pi * k / 8
        bl.append(circle(radius, ds, center=(1.2 * np.cos(th), 1.2 * np.sin(th))))
    bl.append(plate(1.0, ds))
    return concat(*bl)

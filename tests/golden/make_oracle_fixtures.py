"""Generates tests/golden/c1_dirichlet_128.npz: the CPU oracle's outputs on BASELINE config C1
(Dirichlet Poisson on a circle, 128x128 grid, Yang3, examples/dirichlet.ipynb setup scaled to the
(-2,2)^2 synthetic grid of SURVEY.md section 8d), on seeded inputs.

    python tests/golden/make_oracle_fixtures.py

The file is committed.  tests/test_golden.py checks (CPU) that the oracle still reproduces it and
(-m gpu) that the CUDA path matches it, so parity is pinned to a fixed artefact and not only to an
oracle that runs next to the GPU code.  The reference is Julia and cannot run here (SURVEY.md fact
3), so these are oracle outputs; the numbers printed by the reference's own executed notebooks are
in reference_notebook_values.json (extract_reference_goldens.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def build():
    import ilm_b200 as ilm            # host-side helpers only (grid, body, LGF table): no GPU needed
    import ilm_oracle as o

    g = ilm.PhysicalGrid.centered(128)
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    G = ilm.lgf.lgf_table(128)
    oc = o.ScalarCache(o.Grid(g.NX, g.NY, g.dx, g.I0), *body[:5], G)
    N = len(body[0])
    rng = np.random.default_rng(0)
    w = rng.standard_normal(o.field_shape(o.PRIMAL, g.NX, g.NY))
    f = np.random.default_rng(1).standard_normal(N)
    fplus = np.asarray(body[0]).copy()                       # f+ = x on the surface (dirichlet.jl:21,75-77)
    sol_f, sol_s, S = o.dirichlet_solve(oc, fplus)
    tab = oc.tabs[o.PRIMAL]
    out = dict(
        NX=g.NX, NY=g.NY, dx=g.dx, I0=np.asarray(g.I0), x=body[0], y=body[1], nx=body[2], ny=body[3], ds=body[4],
        lgf=G, w=w, f=f,
        table_idx=np.transpose(tab.linear_index(), (0, 2, 1)), table_wR=np.transpose(tab.wR, (0, 2, 1)),
        table_wE=np.transpose(tab.wE, (0, 2, 1)),
        regularize_f=o.regularize(tab, f), interpolate_w=o.interpolate(tab, w),
        inverse_laplacian_w=oc.inverse_laplacian(w.copy()),
        divergence_grad_w=oc.divergence(*oc.grad(w)),
        mask=oc.mask(), S=S, dirichlet_s=sol_s, dirichlet_f=sol_f,
        CLinvCT=oc.create_CLinvCT(), nRTRn=oc.create_nRTRn(), surface_filter=oc.create_surface_filter(),
    )
    return out


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_dirichlet_128.npz")
    np.savez_compressed(dst, **build())
    print("wrote", dst, os.path.getsize(dst), "bytes")

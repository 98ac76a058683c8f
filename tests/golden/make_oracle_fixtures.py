"""Generates tests/golden/c1_dirichlet_128.npz and forcing_helmholtz_64.npz: the CPU oracle's outputs on BASELINE config C1
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


def build_forcing_helmholtz():
    """forcing_helmholtz_64.npz: the oracle's forcing-region contributions (src/forcing.jl:456-494) and Helmholtz
    decomposition (src/helmholtz.jl:84-307) on a 64 x 60 grid with an off-centre ellipse, seeded inputs."""
    import ilm_b200 as ilm
    import ilm_oracle as o

    NX, NY, dx, I0 = 64, 60, 0.07, (31, 29)
    G = ilm.lgf.lgf_table(64)
    og = o.Grid(NX, NY, dx, I0)
    body = ilm.bodies.ellipse(0.9, 0.5, 1.4 * dx, center=(0.1, -0.05))
    shape = ilm.RigidTransform((0.3, 0.2), 0.4)(ilm.bodies.rectangle(0.5, 0.3, 1.4 * dx))
    oc_shape = o.ScalarCache(og, *shape[:5], G)
    vc = o.VectorCache(og, *body[:5], G)
    rng = np.random.default_rng(21)
    T = rng.standard_normal(o.field_shape(o.PRIMAL, NX, NY))
    line_str = rng.standard_normal(oc_shape.N)
    px, py = np.array([-1.2, 0.5, 0.013]), np.array([0.5, 0.5, -0.777])
    pstr = np.array([1.0, -1.0, 2.5])
    m = oc_shape.mask()
    dT = o.forcing_area(np.zeros_like(T), 2.0 * (1.5 - T), m)
    dT = o.forcing_line(dT, oc_shape.tabs[o.PRIMAL], line_str)
    dT = o.forcing_line(dT, o.point_collection_table(og, px, py, o.PRIMAL, "m4prime"), pstr)
    w = rng.standard_normal(o.field_shape(o.DUAL, NX, NY))
    d = rng.standard_normal(o.field_shape(o.PRIMAL, NX, NY))
    dvu, dvv = rng.standard_normal(vc.N), rng.standard_normal(vc.N)
    psi, phi = o.helmholtz_potentials(vc, w, d, dvu, dvv)
    vu, vv = o.vecfield_from_potentials(vc, psi, phi, None)
    return dict(NX=NX, NY=NY, dx=dx, I0=np.asarray(I0), lgf=G, body=np.stack(body[:5]), shape=np.stack(shape[:5]),
                T=T, line_str=line_str, px=px, py=py, pstr=pstr, shape_mask=m, forcing_dT=dT,
                w=w, d=d, dv=np.concatenate([dvu, dvv]), psi=psi, phi=phi, v_u=vu, v_v=vv,
                masked_w=o.helmholtz_jump(vc, "cross", -1, dvu, dvv, w), masked_d=o.helmholtz_jump(vc, "dot", -1, dvu, dvv, d))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("c1_dirichlet_128.npz", build), ("forcing_helmholtz_64.npz", build_forcing_helmholtz)):
        dst = os.path.join(here, name)
        np.savez_compressed(dst, **fn())
        print("wrote", dst, os.path.getsize(dst), "bytes")

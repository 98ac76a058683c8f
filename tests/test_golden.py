"""Committed golden fixtures (tests/golden/):

* reference_notebook_values.json -- numbers printed by the reference's own executed notebooks
  (extract_reference_goldens.py); the oracle must reproduce them (CPU).
* c1_dirichlet_128.npz -- oracle outputs on BASELINE config C1 (make_oracle_fixtures.py); the
  oracle must still reproduce the file (CPU) and the CUDA path must match it (-m gpu): bit-exact
  for tables / regularization / stencils, 1e-12 norm-wise for everything through the FFT.
* forcing_helmholtz_64.npz -- oracle outputs of the forcing regions (src/forcing.jl) and the Helmholtz decomposition
  (src/helmholtz.jl) on a 64 x 60 grid (same script), checked the same two ways."""
import json
import os

import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o
from ilm_b200 import _lib as L

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12


@pytest.fixture(scope="module")
def nbvals():
    return json.load(open(os.path.join(HERE, "reference_notebook_values.json")))


@pytest.fixture(scope="module")
def c1file():
    return np.load(os.path.join(HERE, "c1_dirichlet_128.npz"))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ---------------------------------------------------------------- reference notebook values (CPU)
def test_notebook_circle_and_normals(nbvals):
    """examples/caches.ipynb: Circle(1.0, 1.4*0.01) has 448 points, sum(ds) and the first/last
    printed normal components."""
    npts = int(nbvals["caches_circle"]["numbers"][0])
    x, y, nx, ny, ds = ilm.bodies.circle(1.0, 1.4 * 0.01)[:5]
    assert len(x) == npts == 448
    assert abs(ds.sum() - nbvals["caches_dot_ones_surface"]["values"][0]) < 1e-11
    assert abs(ds.sum() - nbvals["caches_integrate_ones_surface"]["values"][0]) < 1e-11
    vals = nbvals["caches_normals_head"]["values"]          # 10 leading nx, then the trailing ny (after the ellipsis)
    # RigidBodyTools (un-vendored; bodies are INPUTS of the hot path) derives its normals from the
    # discretised shape: they deviate by an alternating +-1.5e-6 rad from the exact midpoint normals
    # (its first normal is cos(1.5e-6), not 1)
    assert np.abs(nx[:10] - np.array(vals[:10])).max() < 2e-6
    tail = np.array(vals[10:])
    assert np.abs(ny[-len(tail):] - tail).max() < 2e-6
    assert len(ilm.bodies.circle(0.5, 1.4 * 0.01)[0]) == int(nbvals["multbodies_circle"]["numbers"][0])


def test_notebook_grid_integrals(nbvals):
    """examples/caches.ipynb: PhysicalGrid((406,406),(203,203),0.01,...); integrate(ones_grid) = 16.3216."""
    nums = nbvals["caches_grid"]["numbers"]
    NX, NY, I0x, I0y, dx = int(nums[1]), int(nums[2]), int(nums[3]), int(nums[4]), nums[5]
    g = o.Grid(NX, NY, dx, (I0x, I0y))
    ones = np.ones(o.field_shape(o.PRIMAL, NX, NY))
    val = o.dot_grid(g, ones, ones, o.PRIMAL)
    assert abs(val - nbvals["caches_integrate_ones_grid"]["values"][0]) < 1e-10
    xs, ys = g.coords(o.PRIMAL)
    assert abs(xs[0] - nums[6]) < 1e-12 and abs(xs[-1] - nums[7]) < 1e-12


def test_notebook_layers_dot(nbvals):
    """examples/Layers.ipynb cell 64: dot(qx, Rf*nrm, g) on the 600^2 grid, dx = 0.02, circle R = 1,
    ds = 1.5 dx (cells 5, 19); the regularization weights are dlengthmid(body) (cell 19)."""
    dx = 0.02
    g = o.Grid(600, 600, dx, (300, 300))
    x, y, nx, ny, ds = ilm.bodies.circle(1.0, 1.5 * dx)[:5]
    N = len(x)
    dlmid = np.full(N, np.sin(2 * np.pi / N))            # 0.5*|x[k+1]-x[k-1]| on the unit circle
    q = o.regularize(o.build_table(g, x, y, dlmid, o.XEDGE), nx)
    xu = g.coords(o.XEDGE)[0]
    got = o.dot_grid(g, np.broadcast_to(xu[:, None], q.shape), q, o.XEDGE)
    ref = nbvals["layers_dot_x_Rf_n"]["values"][0]
    assert abs(got - ref) < 1e-13 * abs(ref)


# ---------------------------------------------------------------- oracle vs the committed C1 file (CPU)
def test_oracle_reproduces_c1_fixture(c1file):
    d = c1file
    g = o.Grid(int(d["NX"]), int(d["NY"]), float(d["dx"]), tuple(int(v) for v in d["I0"]))
    oc = o.ScalarCache(g, d["x"], d["y"], d["nx"], d["ny"], d["ds"], d["lgf"])
    tab = oc.tabs[o.PRIMAL]
    assert np.array_equal(np.transpose(tab.linear_index(), (0, 2, 1)), d["table_idx"])
    assert np.array_equal(np.transpose(tab.wR, (0, 2, 1)), d["table_wR"])
    assert np.array_equal(o.regularize(tab, d["f"]), d["regularize_f"])
    assert relerr(o.interpolate(tab, d["w"]), d["interpolate_w"]) < 1e-14
    assert relerr(oc.inverse_laplacian(d["w"].copy()), d["inverse_laplacian_w"]) < 1e-13
    assert relerr(oc.mask(), d["mask"]) < 1e-13
    f, s, S = o.dirichlet_solve(oc, np.asarray(d["x"]).copy())
    assert relerr(S, d["S"]) < 1e-13
    assert relerr(f, d["dirichlet_f"]) < 1e-11
    assert np.array_equal(ilm.lgf.lgf_table(128), d["lgf"])


# ---------------------------------------------------------------- CUDA path vs the committed C1 file
@pytest.mark.gpu
def test_gpu_matches_c1_fixture(c1file):
    d = c1file
    g = ilm.PhysicalGrid(int(d["NX"]), int(d["NY"]), float(d["dx"]), tuple(int(v) for v in d["I0"]))
    body = (d["x"], d["y"], d["nx"], d["ny"], d["ds"])
    cache = ilm.SurfaceScalarCache(body, g, lgf_table=d["lgf"])
    idx, wR, wE = cache.table(L.NODES_PRIMAL)
    assert np.array_equal(idx, d["table_idx"])
    assert np.array_equal(wR, d["table_wR"]) and np.array_equal(wE, d["table_wE"])
    fd = ilm.ScalarData(cache.N, data=d["f"].copy())
    s = cache.zeros_grid()
    ilm.regularize(s, fd, cache)
    assert np.array_equal(s.array(), d["regularize_f"])
    w = cache.zeros_grid().set(d["w"])
    out = cache.zeros_surface()
    ilm.interpolate(out, w, cache)
    assert relerr(out.data, d["interpolate_w"]) < RTOL
    q = cache.zeros_gridgrad()
    ilm.grad(q, w, cache)
    p = cache.zeros_grid()
    ilm.divergence(p, q, cache)
    assert np.array_equal(p.array(), d["divergence_grad_w"])
    ilm.inverse_laplacian(w, cache)
    assert relerr(w.array(), d["inverse_laplacian_w"]) < RTOL
    assert relerr(ilm.mask(cache).array(), d["mask"]) < RTOL
    S = ilm.create_RTLinvR(cache)
    assert relerr(S, d["S"]) < RTOL
    assert relerr(ilm.create_CLinvCT(cache), d["CLinvCT"]) < RTOL
    assert relerr(ilm.create_nRTRn(cache), d["nRTRn"]) < RTOL
    assert relerr(ilm.create_surface_filter(cache), d["surface_filter"]) < RTOL
    f, sm, _ = ilm.dirichlet_poisson(cache, np.asarray(d["x"]).copy())
    cond = np.linalg.cond(d["S"])
    assert relerr(sm.data, d["dirichlet_s"]) < 50 * cond * np.finfo(float).eps
    assert relerr(f.array(), d["dirichlet_f"]) < 1e-10


# ---------------------------------------------------------------- forcing regions / Helmholtz fixture
@pytest.fixture(scope="module")
def fhfile():
    return np.load(os.path.join(HERE, "forcing_helmholtz_64.npz"))


def _fh_oracle(d):
    g = o.Grid(int(d["NX"]), int(d["NY"]), float(d["dx"]), tuple(int(v) for v in d["I0"]))
    return g, o.ScalarCache(g, *d["shape"], d["lgf"]), o.VectorCache(g, *d["body"], d["lgf"])


def test_oracle_reproduces_forcing_helmholtz_fixture(fhfile):
    d = fhfile
    g, ocs, vc = _fh_oracle(d)
    m = ocs.mask()
    assert relerr(m, d["shape_mask"]) < 1e-13
    dT = o.forcing_area(np.zeros_like(d["T"]), 2.0 * (1.5 - d["T"]), d["shape_mask"])
    dT = o.forcing_line(dT, ocs.tabs[o.PRIMAL], d["line_str"])
    dT = o.forcing_line(dT, o.point_collection_table(g, d["px"], d["py"], o.PRIMAL, "m4prime"), d["pstr"])
    assert np.array_equal(dT, d["forcing_dT"])
    n = vc.N
    psi, phi = o.helmholtz_potentials(vc, d["w"], d["d"], d["dv"][:n], d["dv"][n:])
    assert relerr(psi, d["psi"]) < 1e-13 and relerr(phi, d["phi"]) < 1e-13
    vu, vv = o.vecfield_from_potentials(vc, d["psi"], d["phi"], None)
    assert np.array_equal(vu, d["v_u"]) and np.array_equal(vv, d["v_v"])
    assert np.array_equal(o.helmholtz_jump(vc, "cross", -1, d["dv"][:n], d["dv"][n:], d["w"]), d["masked_w"])


@pytest.mark.gpu
def test_gpu_matches_forcing_helmholtz_fixture(fhfile):
    d = fhfile
    g = ilm.PhysicalGrid(int(d["NX"]), int(d["NY"]), float(d["dx"]), tuple(int(v) for v in d["I0"]))
    base = ilm.SurfaceScalarCache(tuple(d["body"]), g, lgf_table=d["lgf"])
    T = base.zeros_grid().set(d["T"])

    def area(sig, TT, t, fr, pp):
        sig.set(2.0 * (1.5 - TT.numpy()))

    def line(sig, TT, t, fr, pp):
        sig.set(d["line_str"])

    def point(sig, TT, t, fr, pp):
        sig.set(d["pstr"])

    shape = tuple(d["shape"])
    fc = ilm.ForcingModelAndRegion([ilm.AreaForcingModel(shape, ilm.RigidTransform(), area),
                                    ilm.LineForcingModel(shape, ilm.RigidTransform(), line),
                                    ilm.PointForcingModel((d["px"], d["py"]), point, ddftype="m4prime")], base)
    dT = base.zeros_grid()
    ilm.apply_forcing(dT, T, None, 0.0, fc, None, None, base)
    assert relerr(fc[0].region_cache.mask.array(), d["shape_mask"]) < RTOL
    assert relerr(dT.array(), d["forcing_dT"]) < RTOL
    vc = ilm.SurfaceVectorCache(tuple(d["body"]), g, parent=base)
    w, dd, dv = vc.zeros_gridcurl().set(d["w"]), vc.zeros_griddiv().set(d["d"]), vc.zeros_surface().set(d["dv"])
    psi, phi, v = vc.zeros_gridcurl(), vc.zeros_griddiv(), vc.zeros_grid()
    ilm.potentials_from_masked_fields(psi, phi, w, dd, dv, vc)
    assert relerr(psi.array(), d["psi"]) < RTOL and relerr(phi.array(), d["phi"]) < RTOL
    ilm.vecfield_helmholtz(v, w, dd, dv, None, vc)
    assert relerr(v.u, d["v_u"]) < RTOL and relerr(v.v, d["v_v"]) < RTOL
    mw, md = vc.zeros_gridcurl(), vc.zeros_griddiv()
    ilm.masked_curlv_from_curlv_masked(mw, w, dv, vc)
    ilm.masked_divv_from_divv_masked(md, dd, dv, vc)
    assert np.array_equal(mw.array(), d["masked_w"]) and np.array_equal(md.array(), d["masked_d"])


# ---------------------------------------------------------------- full-problem values printed by the reference's notebooks
# These pin the oracle (and through it the CUDA path) to the REFERENCE ITSELF at 1e-14: a 54-step IF-HERK time
# march (integrator of the un-vendored ConstrainedSystems.jl, plan_intfact, DDF tables, surface operators) and a
# two-body Neumann solve (LGF inverse Laplacian, create_CLinvCT, surface grad / divergence / curl, dense solve).
NB_GRID = (408, 0.01, (204, 204))          # PhysicalGrid((-2,2),(-2,2),0.01) of the notebooks: Nodes{Primal,408,408}


def _nb_node(x, y):
    NX, dx, I0 = NB_GRID
    return I0[0] - 1 + int(round(x / dx)), I0[1] - 1 + int(round(y / dx))          # 0-based primal-node index


def test_notebook_heatconduction_time_marching(nbvals):
    """examples/heatconduction.ipynb cells 19-62: DirichletHeatConduction on Circle(1, 1.4 dx), T+ = 0, T- = 1,
    kappa = Fo = 1, LiskaIFHERK; `Tfcn(-0.9,0)` at t = 0.0051 and t = 0.0054 (a grid node, so the cubic-spline
    evaluation returns the nodal value)."""
    from ilm_b200 import timemarching as tm
    NX, dx, I0 = NB_GRID
    assert nbvals["heatconduction_state_size"]["numbers"][:2] == [408.0, 408.0]
    dt = nbvals["heatconduction_dt"]["values"][0]
    assert abs(dt - dx ** 2) < 1e-18
    g = o.Grid(NX, NX, dx, I0)
    body = ilm.bodies.circle(1.0, 1.4 * dx)
    oc = o.ScalarCache(g, *body[:5], ilm.lgf.lgf_table(NX), workers=os.cpu_count() or 1)
    tab = tm.LISKA_IFHERK
    tables = {a: ilm.lgf.intfact_table(a, NX) for a in (0.0, 0.5)}
    T = np.zeros(o.field_shape(o.PRIMAL, NX, NX))
    i, j = _nb_node(-0.9, 0.0)
    reuse, got = {}, {}
    for n in range(1, 55):
        T, _ = o.heat_ifherk_step(oc, T, (n - 1) * dt, dt, 1.0, tab["a"], tab["c"], tables, 0.0, 1.0, reuse=reuse)
        got[n] = T[i, j]
    r51, r54 = nbvals["heatconduction_T_t0051"]["values"][0], nbvals["heatconduction_T_t0054"]["values"][0]
    assert abs(got[51] - r51) < 1e-12 * abs(r51), (got[51], r51)
    assert abs(got[54] - r54) < 1e-12 * abs(r54), (got[54], r54)


def _neumann_two_bodies():
    NX, dx, I0 = NB_GRID
    sq, ci = ilm.bodies.rectangle(1.0, 1.0, 1.4 * dx), ilm.bodies.circle(0.25, 1.4 * dx)       # Square(1.0, ds), Circle(0.25, ds)
    body = ilm.bodies.concat(sq, ci)
    n1 = len(sq[0])
    vnp = np.zeros(len(body[0]))
    vnp[n1:] = body[2][n1:]                       # v_n+ = n_x on body 2 (copyto!(vnplus, nrm.u, base_cache, 2))
    return body, n1, vnp


def _added_mass(df, body, n1):
    """M = -integrate(df o nrm, sys, 2) (examples/neumann.ipynb cell 37)."""
    df = np.asarray(df)
    return -np.array([np.sum(df[n1:] * body[2][n1:] * body[4][n1:]), np.sum(df[n1:] * body[3][n1:] * body[4][n1:])])


def test_notebook_neumann_added_mass(nbvals):
    """examples/neumann.ipynb cells 27-37: a circle of radius 1/4 moving inside a square box of half-side 1."""
    NX, dx, I0 = NB_GRID
    body, n1, vnp = _neumann_two_bodies()
    oc = o.ScalarCache(o.Grid(NX, NX, dx, I0), *body[:5], ilm.lgf.lgf_table(NX), workers=os.cpu_count() or 1)
    M = _added_mass(o.neumann_solve(oc, vnp)[1], body, n1)
    ref = nbvals["neumann_added_mass"]["values"]
    assert abs(M[0] - ref[0]) < 1e-12 * abs(ref[0]), (M, ref)
    # the reference's own y-component is -4.4e-10 (its body normals deviate by 1.5e-6 rad, see above); ours is symmetric
    assert abs(M[1]) < 1e-12 and abs(M[1] - ref[1]) < 1e-9


def test_notebook_multbodies_volume(nbvals):
    """examples/multbodies.ipynb cell 16: integrate(pointwise_dot(pts, nrm), cache, 3) = 2 x area of Circle(0.5, 1.4 dx)
    placed at (-1, 1)."""
    x, y, nx, ny, ds = ilm.bodies.circle(0.5, 1.4 * 0.01, center=(-1.0, 1.0))
    assert abs(np.sum((x * nx + y * ny) * ds) - nbvals["multbodies_volume"]["values"][0]) < 1e-12


def test_notebook_circle_endpoints(nbvals):
    """examples/heatconduction.ipynb cell 69: maxvelocity of Circle(1, 1.4 dx) under the deformation u = xy/4, v = (x^2-y^2)/4
    is 0.24998770586539437 at point 225.  RigidBodyTools averages the velocities of a segment's END points; with the
    midpoints ON the circle at theta_k = 2 pi k / N (ilm.bodies.circle) the end points sit at radius 1/cos(pi/N), which
    gives (1 - tan^2(pi/N))/4 at theta = pi -- the circumscribed-polygon convention that also reproduces sum(ds)."""
    nums = nbvals["heatconduction_maxvelocity"]["numbers"]
    x, y, nx, ny, ds = ilm.bodies.circle(1.0, 1.4 * 0.01)
    n = len(x)
    th = np.arctan2(y, x)
    rho, a = 1.0 / np.cos(np.pi / n), np.pi / n
    xe = rho * np.cos(th[:, None] + np.array([-a, a])[None, :])
    ye = rho * np.sin(th[:, None] + np.array([-a, a])[None, :])
    u, v = (0.25 * xe * ye).mean(axis=1), (0.25 * (xe ** 2 - ye ** 2)).mean(axis=1)
    speed = np.hypot(u, v)
    k = int(nums[1]) - 1
    assert abs(speed[k] - nums[0]) < 1e-12 and abs(speed.max() - nums[0]) < 1e-12
    assert abs(x[k] + 1.0) < 1e-12 and abs(y[k]) < 1e-12

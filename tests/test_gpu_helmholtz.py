"""GPU parity of the Helmholtz decomposition (src/helmholtz.jl) against the oracle, plus the reference's own
round-trip check (test/surface_ops.jl:319-372).  Potentials pass through the FFT (1e-12 norm-wise); the jump
conversions and the recomposition v = curl psi + grad phi + vp are bit-exact given their inputs."""
import numpy as np
import pytest

import ilm_b200 as ilm
import ilm_oracle as o

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module", params=[False, True], ids=["host", "device"])
def case(request):
    NX, NY, dx, I0 = 104, 96, 0.04, (52, 47)
    g = ilm.PhysicalGrid(NX, NY, dx, I0)
    G = ilm.lgf.lgf_table(max(NX, NY))
    body = ilm.bodies.ellipse(1.0, 0.6, 1.4 * dx, center=(0.05, -0.1))
    vcache = ilm.SurfaceVectorCache(body, g, lgf_table=G, device=request.param)
    oc = o.VectorCache(o.Grid(NX, NY, dx, I0), *body[:5], G)
    rng = np.random.default_rng(11)
    w, d, dv = vcache.zeros_gridcurl(), vcache.zeros_griddiv(), vcache.zeros_surface()
    w.set(rng.standard_normal(len(w)))
    d.set(rng.standard_normal(len(d)))
    dv.set(rng.standard_normal(len(dv)))
    return g, G, vcache, oc, w, d, dv


def test_potentials_and_field_match_oracle(case):
    g, G, vc, oc, w, d, dv = case
    psi_r, phi_r = o.helmholtz_potentials(oc, w.array(), d.array(), dv.u, dv.v)
    psi, phi = vc.zeros_gridcurl(), vc.zeros_griddiv()
    ilm.vectorpotential_from_masked_curlv(psi, w, dv, vc, ilm.VectorPotentialCache(vc))
    ilm.scalarpotential_from_masked_divv(phi, d, dv, vc, ilm.ScalarPotentialCache(vc))
    assert relerr(psi.array(), psi_r) < 1e-12 and relerr(phi.array(), phi_r) < 1e-12
    # both in one transform: same values as the two separate solves up to the transform's rounding
    psi2, phi2 = vc.zeros_gridcurl(), vc.zeros_griddiv()
    ilm.potentials_from_masked_fields(psi2, phi2, w, d, dv, vc)
    assert relerr(psi2.array(), psi_r) < 1e-12 and relerr(phi2.array(), phi_r) < 1e-12
    # potentials of plain fields (no jump)
    ilm.vectorpotential_from_curlv(psi, w, vc)
    ilm.scalarpotential_from_divv(phi, d, vc)
    assert relerr(psi.array(), oc.inverse_laplacian(w.array()) * -1.0) < 1e-12
    assert relerr(phi.array(), oc.inverse_laplacian(d.array())) < 1e-12

    vp = vc.zeros_grid()
    vp.set(np.random.default_rng(12).standard_normal(len(vp)))
    v = vc.zeros_grid()
    v.fill(3.0)
    ilm.vecfield_helmholtz(v, w, d, dv, vp, vc, ilm.VectorFieldCache(vc))
    ur, vr = o.vecfield_helmholtz(oc, w.array(), d.array(), dv.u, dv.v, (vp.u, vp.v))
    assert relerr(v.u, ur) < 1e-12 and relerr(v.v, vr) < 1e-12
    # recomposition is bit-exact given the library's own potentials
    v2 = vc.zeros_grid()
    ilm.helmholtz.L.check(vc._lib.ilm_vecfield_from_potentials(vc._plan, ilm.api._ptr(psi2.data), ilm.api._ptr(phi2.data),
                                                               ilm.api._ptr(vp.data), ilm.api._ptr(v2.data)))
    u3, v3 = o.vecfield_from_potentials(oc, psi2.array(), phi2.array(), (vp.u, vp.v))
    assert np.array_equal(v2.u, u3) and np.array_equal(v2.v, v3)
    assert np.array_equal(v2.numpy(), v.numpy())          # the fused entry point is the composition of the two
    # single-potential fields
    ilm.vecfield_from_vectorpotential(v2, psi2, vc)
    cu, cv = oc.curl_n2e(psi2.array())
    assert np.array_equal(v2.u, cu) and np.array_equal(v2.v, cv)
    ilm.vecfield_from_scalarpotential(v2, phi2, vc)
    gu, gv = oc.grad(phi2.array())
    assert np.array_equal(v2.u, gu) and np.array_equal(v2.v, gv)


def test_jump_conversions_bit_exact(case):
    g, G, vc, oc, w, d, dv = case
    out = vc.zeros_gridcurl()
    ilm.masked_curlv_from_curlv_masked(out, w, dv, vc)
    assert np.array_equal(out.array(), o.helmholtz_jump(oc, "cross", -1, dv.u, dv.v, w.array()))
    ilm.curlv_masked_from_masked_curlv(out, w, dv, vc)
    assert np.array_equal(out.array(), o.helmholtz_jump(oc, "cross", +1, dv.u, dv.v, w.array()))
    outp = vc.zeros_griddiv()
    ilm.masked_divv_from_divv_masked(outp, d, dv, vc)
    assert np.array_equal(outp.array(), o.helmholtz_jump(oc, "dot", -1, dv.u, dv.v, d.array()))
    ilm.divv_masked_from_masked_divv(outp, d, dv, vc)
    assert np.array_equal(outp.array(), o.helmholtz_jump(oc, "dot", +1, dv.u, dv.v, d.array()))
    # in place (out aliases the input field)
    w2 = vc.zeros_gridcurl()
    w2.set(w.numpy())
    ilm.masked_curlv_from_curlv_masked(w2, w2, dv, vc)
    assert np.array_equal(w2.array(), o.helmholtz_jump(oc, "cross", -1, dv.u, dv.v, w.array()))
    with pytest.raises(ilm.MethodError):
        ilm.masked_curlv_from_curlv_masked(outp, w, dv, vc)
    with pytest.raises(ilm.DimensionMismatch):
        ilm.masked_curlv_from_curlv_masked(out, w, ilm.VectorData(3), vc)


def test_reference_round_trip(case):
    """test/surface_ops.jl:336-350: curl and divergence of the recomposed field, un-jumped, return w and d (1e-8)."""
    g, G, vc, oc, w, d, dv = case
    v = vc.zeros_grid()
    ilm.vecfield_helmholtz(v, w, d, dv, (0.0, 0.0), vc, ilm.VectorFieldCache(vc))
    w2, mw = vc.zeros_gridcurl(), vc.zeros_gridcurl()
    ilm.curl(w2, v, vc)
    ilm.masked_curlv_from_curlv_masked(mw, w2, dv, vc)
    assert np.abs(mw.array()[1:-1, 1:-1] - w.array()[1:-1, 1:-1]).max() < 1e-8
    d2, md = vc.zeros_griddiv(), vc.zeros_griddiv()
    ilm.divergence(d2, v, vc)
    ilm.masked_divv_from_divv_masked(md, d2, dv, vc)
    assert np.abs(md.array()[1:-1, 1:-1] - d.array()[1:-1, 1:-1]).max() < 1e-8


def test_no_immersed_points_and_uniform_field(case):
    """test/surface_ops.jl:355-371: a cache without points gives zero potentials and field for zero inputs; the
    potentials of a uniform field recompose to it (src/helmholtz.jl:317-348)."""
    g, G, vc, oc, w, d, dv = case
    z = np.zeros(0)
    vc0 = ilm.SurfaceVectorCache((z, z, z, z, z), g, lgf_table=G, device=vc.device)
    w0, psi, d0, phi = vc0.zeros_gridcurl(), vc0.zeros_gridcurl(), vc0.zeros_griddiv(), vc0.zeros_griddiv()
    v, dv0 = vc0.zeros_grid(), vc0.zeros_surface()
    ilm.vectorpotential_from_masked_curlv(psi, w0, dv0, vc0)
    ilm.scalarpotential_from_masked_divv(phi, d0, dv0, vc0)
    ilm.vecfield_helmholtz(v, w0, d0, dv0, (0.0, 0.0), vc0)
    assert np.abs(psi.numpy()).max() == 0.0 and np.abs(phi.numpy()).max() == 0.0 and np.abs(v.numpy()).max() == 0.0
    out = vc0.zeros_gridcurl()
    ilm.masked_curlv_from_curlv_masked(out, w, dv0, vc0)
    assert np.array_equal(out.numpy(), w.numpy())
    # uniform field (Vx, Vy) = (1.5, -0.5) from either potential
    ilm.vectorpotential_uniformvecfield(psi, 1.5, -0.5, vc0)
    ilm.vecfield_from_vectorpotential(v, psi, vc0)
    assert np.abs(v.u - 1.5).max() < 1e-12 and np.abs(v.v + 0.5).max() < 1e-12
    ilm.scalarpotential_uniformvecfield(phi, 1.5, -0.5, vc0)
    ilm.vecfield_from_scalarpotential(v, phi, vc0)
    assert np.abs(v.u[1:-1, :] - 1.5).max() < 1e-12 and np.abs(v.v[:, 1:-1] + 0.5).max() < 1e-12
    ilm.vecfield_uniformvecfield(v, 2.0, 3.0, vc0)
    assert np.array_equal(v.u, np.full_like(v.u, 2.0)) and np.array_equal(v.v, np.full_like(v.v, 3.0))

# ImmersedLayersB200.jl -- thin `ccall` shim that routes the operator hot path of ImmersedLayers.jl
# (regularize!/interpolate!, staggered stencils, inverse_laplacian!, Schur builders, mask, convective
# terms) to libilm_b200.so (include/ilm_b200.h).  The problem / cache / system framework of
# ImmersedLayers stays untouched: a `B200Cache` wraps a `BasicILMCache` plus the device plan, and every
# method below has the signature of the reference method it replaces (file:line cited).
#
# NOT EXERCISED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain (SURVEY.md fact 3).  The
# Python mirror (immersedlayers.jl_b200/api.py) binds the same symbols and is what the tests drive; this
# file is the maintainer-side binding that INTEGRATION.md describes, kept next to the ABI it targets.
module ImmersedLayersB200

using ImmersedLayers, CartesianGrids, Libdl
import ImmersedLayers: regularize!, interpolate!, regularize_normal!, normal_interpolate!,
    regularize_normal_cross!, normal_cross_interpolate!, surface_divergence!, surface_grad!, surface_curl!,
    surface_divergence_cross!, surface_grad_cross!, surface_curl_cross!, inverse_laplacian!, mask!,
    complementary_mask!, create_RTLinvR, create_CLinvCT, create_GLinvD, create_GLinvD_cross, create_nRTRn,
    create_surface_filter, convective_derivative!, w_cross_v!
import CartesianGrids: divergence!, grad!, curl!, laplacian!

export B200Cache, update_points!, lu_b200, solve_b200!, intfact_kernel, convolve!, create_RTHR

const lib = get(ENV, "ILM_B200_LIB", "libilm_b200.so")

struct IlmGrid            # mirrors `ilm_grid`
    NX::Cint; NY::Cint; dx::Cdouble; I0x::Cint; I0y::Cint
end

# layouts / modes of include/ilm_b200.h
const NODES_PRIMAL, NODES_DUAL, XEDGES, YEDGES, EDGES, EDGEGRAD = Cint.(0:5)
const NORMAL, CROSS = Cint(0), Cint(1)
const RTLINVR, CLINVCT, GLINVD, GLINVD_CROSS = Cint.(0:3)
const DDF = Dict(CartesianGrids.Yang3 => 0, CartesianGrids.M3 => 1, CartesianGrids.Roma => 2,
                 CartesianGrids.M4prime => 3, CartesianGrids.Witchhat => 4)

function check(status::Integer)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:ilm_last_error, lib), Cstring, ()))
    status == 2 && throw(DimensionMismatch(msg))      # ILM_ESIZE  <-> AssertionError / DimensionMismatch
    status == 1 && throw(ArgumentError(msg))          # ILM_EINVAL <-> MethodError / ArgumentError
    error("ilm_b200 (status $status): $msg")
end

"""
    B200Cache(cache::BasicILMCache; ddftype=CartesianGrids.Yang3)

Device plan for `cache` (what `_surfacecache`, `_get_regularization`, `_get_laplacian` build,
src/cache.jl:232-261,305-324).  The LGF table handed over is CartesianGrids' own `LGF_TABLE`, so the
multiplier on the device is built from the numbers the CPU path uses.
"""
mutable struct B200Cache{N,SCA,BC<:BasicILMCache{N,SCA}}
    base::BC
    plan::Ptr{Cvoid}
    vector::Bool
    parent::Any       # the cache whose Laplacian a shared plan aliases (ilm_plan_create_shared): finalizers run in
                      # arbitrary order, so the child keeps its parent reachable for as long as it lives; nothing otherwise
end
B200Cache{N,SCA,BC}(base::BC, plan::Ptr{Cvoid}, vector::Bool) where {N,SCA,BC<:BasicILMCache{N,SCA}} =
    B200Cache{N,SCA,BC}(base, plan, vector, nothing)

# LGF table with at least n x n entries.  CartesianGrids ships LGF_TABLE (1600 x 1600, SURVEY.md fact 7); beyond it the
# entries come from the two-term far-field expansion (ln r + gamma + ln(8)/2)/2pi - cos(4 theta)/(24 pi r^2), accurate to
# < 1e-13 for r >= 800 (SURVEY.md section 8c), which is what lets a 4096^2 (or 16384^2) grid be planned at all.
const DEFAULT_DDF = CartesianGrids.Yang3
function lgf_table(n::Integer)
    tbl = CartesianGrids.LGF_TABLE
    m = size(tbl, 1)
    n <= m && return tbl
    G = Matrix{Float64}(undef, n, n)
    G[1:m, 1:m] .= tbl
    for j in 1:n, i in 1:n
        (i <= m && j <= m) && continue
        x, y = i - 1, j - 1
        r2 = x^2 + y^2
        G[i, j] = (log(r2) / 2 + MathConstants.γ + log(8) / 2) / 2π - cos(4 * atan(y, x)) / (24π * r2)
    end
    return G
end

layout(::Nodes{Primal}) = NODES_PRIMAL
layout(::Nodes{Dual}) = NODES_DUAL
layout(::XEdges{Primal}) = XEDGES
layout(::YEdges{Primal}) = YEDGES
layout(::Edges{Primal}) = EDGES
layout(::EdgeGradient{Primal}) = EDGEGRAD

function B200Cache(cache::BasicILMCache{N,SCA}; ddftype=CartesianGrids.Yang3) where {N,SCA}
    g = cache.g
    NX, NY = size(g); I0 = origin(g); dx = cellsize(g)
    grid = Ref(IlmGrid(NX, NY, dx, I0[1], I0[2]))
    pts, nrm, ds = points(cache), normals(cache), areas(cache)
    tbl = lgf_table(max(NX, NY))                       # ilm_plan_create needs at least max(NX, NY) entries per direction
    factor = SCA == GridScaling ? 1 / dx^2 : 1.0       # src/cache.jl:321-324
    c0 = (MathConstants.γ + log(8) / 2 - log(dx)) / 2π  # far-field constant of L\w (DESIGN.md section 3: unpinned)
    plan = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ilm_plan_create, lib), Cint,
        (Ref{IlmGrid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
         Cint, Cint, Ptr{Cdouble}, Cint, Cdouble, Cdouble, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
        grid, N, pts.u, pts.v, nrm.u, nrm.v, ds.data, DDF[ddftype], SCA == GridScaling ? 0 : 1,
        tbl, size(tbl, 1), c0, factor, C_NULL, plan))
    c = B200Cache{N,SCA,typeof(cache)}(cache, plan[], cache.gdata_cache isa Edges)
    finalizer(x -> ccall((:ilm_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), x.plan), c)
    return c
end

# update_system keeping L (src/system.jl:26-50): new points, same multiplier
function update_points!(c::B200Cache, cache::BasicILMCache{N}) where {N}
    pts, nrm, ds = points(cache), normals(cache), areas(cache)
    check(ccall((:ilm_plan_update_points, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        c.plan, N, pts.u, pts.v, nrm.u, nrm.v, ds.data))
    c.base = cache
    return c
end

const PD = Ptr{Cdouble}
_c2(sym, c, a, b) = check(ccall((sym, lib), Cint, (Ptr{Cvoid}, PD, PD), c.plan, a, b))
_c3(sym, c, k, a, b) = check(ccall((sym, lib), Cint, (Ptr{Cvoid}, Cint, PD, PD), c.plan, k, a, b))

# ---- src/surface_operators.jl:13-88
regularize!(s::Nodes, f::ScalarData, c::B200Cache) = (_c3(:ilm_regularize, c, layout(s), f.data, s.data); s)
regularize!(q::Edges{Primal}, f::VectorData, c::B200Cache) = (_c3(:ilm_regularize, c, EDGES, f.data, q.data); q)
interpolate!(f::ScalarData, s::Nodes, c::B200Cache) = (_c3(:ilm_interpolate, c, layout(s), s.data, f.data); f)
interpolate!(f::VectorData, q::Edges{Primal}, c::B200Cache) = (_c3(:ilm_interpolate, c, EDGES, q.data, f.data); f)
# ---- :98-103, 150-162, 228-233, 270-291 (scalar-cache forms)
regularize_normal!(q::Edges{Primal}, f::ScalarData, c::B200Cache) = (_c3(:ilm_regularize_normal, c, NORMAL, f.data, q.data); q)
regularize_normal_cross!(q::Edges{Primal}, f::ScalarData, c::B200Cache) = (_c3(:ilm_regularize_normal, c, CROSS, f.data, q.data); q)
normal_interpolate!(f::ScalarData, q::Edges{Primal}, c::B200Cache) = (_c3(:ilm_normal_interpolate, c, NORMAL, q.data, f.data); f)
normal_cross_interpolate!(f::ScalarData, q::Edges{Primal}, c::B200Cache) = (_c3(:ilm_normal_interpolate, c, CROSS, q.data, f.data); f)
# ---- vector-cache forms (:113-138, 242-267)
regularize_normal!(q::EdgeGradient{Primal}, v::VectorData, c::B200Cache) = (_c3(:ilm_regularize_normal_tensor, c, Cint(0), v.data, q.data); q)
normal_interpolate!(v::VectorData, q::EdgeGradient{Primal}, c::B200Cache) = (_c3(:ilm_normal_interpolate_tensor, c, Cint(0), q.data, v.data); v)

# ---- src/grid_operators.jl:25-134
divergence!(p::Nodes{Primal}, q::Edges{Primal}, c::B200Cache) = (_c2(:ilm_divergence, c, q.data, p.data); p)
grad!(q::Edges{Primal}, p::Nodes{Primal}, c::B200Cache) = (_c2(:ilm_grad, c, p.data, q.data); q)
curl!(q::Edges{Primal}, s::Nodes{Dual}, c::B200Cache) = (_c2(:ilm_curl_n2e, c, s.data, q.data); q)
curl!(w::Nodes{Dual}, q::Edges{Primal}, c::B200Cache) = (_c2(:ilm_curl_e2n, c, q.data, w.data); w)
grad!(t::EdgeGradient{Primal}, q::Edges{Primal}, c::B200Cache) = (_c2(:ilm_grad_tensor, c, q.data, t.data); t)
divergence!(q::Edges{Primal}, t::EdgeGradient{Primal}, c::B200Cache) = (_c2(:ilm_divergence_tensor, c, t.data, q.data); q)
laplacian!(w::T, s::T, c::B200Cache) where {T<:Nodes} = (_c3(:ilm_laplacian, c, layout(s), s.data, w.data); w)

# ---- :153-179  inverse_laplacian!(w, cache)   (in place; Edges = u and v in one complex transform)
inverse_laplacian!(w::Union{Nodes,XEdges{Primal},YEdges{Primal},Edges{Primal}}, c::B200Cache) =
    (check(ccall((:ilm_inverse_laplacian, lib), Cint, (Ptr{Cvoid}, Cint, PD), c.plan, layout(w), w.data)); w)

# ---- composites (src/surface_operators.jl:357-725)
surface_divergence!(θ::Nodes{Primal}, f::ScalarData, c::B200Cache) = (_c3(:ilm_surface_divergence, c, NORMAL, f.data, θ.data); θ)
surface_divergence_cross!(θ::Nodes{Primal}, f::ScalarData, c::B200Cache) = (_c3(:ilm_surface_divergence, c, CROSS, f.data, θ.data); θ)
surface_grad!(f::ScalarData, ϕ::Nodes{Primal}, c::B200Cache) = (_c3(:ilm_surface_grad, c, NORMAL, ϕ.data, f.data); f)
surface_grad_cross!(f::ScalarData, ϕ::Nodes{Primal}, c::B200Cache) = (_c3(:ilm_surface_grad, c, CROSS, ϕ.data, f.data); f)
surface_curl!(w::Nodes{Dual}, f::ScalarData, c::B200Cache) = (_c3(:ilm_surface_curl_s2n, c, NORMAL, f.data, w.data); w)
surface_curl!(f::ScalarData, s::Nodes{Dual}, c::B200Cache) = (_c3(:ilm_surface_curl_n2s, c, NORMAL, s.data, f.data); f)
surface_curl_cross!(w::Nodes{Dual}, f::ScalarData, c::B200Cache) = (_c3(:ilm_surface_curl_s2n, c, CROSS, f.data, w.data); w)
surface_curl_cross!(f::ScalarData, s::Nodes{Dual}, c::B200Cache) = (_c3(:ilm_surface_curl_n2s, c, CROSS, s.data, f.data); f)
surface_divergence!(q::Edges{Primal}, v::VectorData, c::B200Cache) = (_c3(:ilm_vsurface_divergence, c, Cint(0), v.data, q.data); q)
surface_grad!(v::VectorData, q::Edges{Primal}, c::B200Cache) = (_c3(:ilm_vsurface_grad, c, Cint(0), q.data, v.data); v)
surface_curl!(w::Nodes{Dual}, v::VectorData, c::B200Cache) = (_c2(:ilm_vsurface_curl_s2n, c, v.data, w.data); w)
surface_curl!(v::VectorData, s::Nodes{Dual}, c::B200Cache) = (_c2(:ilm_vsurface_curl_n2s, c, s.data, v.data); v)

# ---- masks (:788-823, 880-923)
_mask!(w, c::B200Cache, comp) = (check(ccall((:ilm_mask_product, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, PD),
                                            c.plan, c.vector ? 1 : 0, layout(w), comp, w.data)); w)
mask!(w::GridData, c::B200Cache) = _mask!(w, c, 0)
complementary_mask!(w::GridData, c::B200Cache) = _mask!(w, c, 1)

# ---- src/matrix_operators.jl
function _schur(c::B200Cache{N}, which, scale) where {N}
    A = Matrix{Float64}(undef, N, N)
    check(ccall((:ilm_create_schur, lib), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cint, Cint, PD), c.plan, which, scale, 0, N, A))
    return A
end
create_RTLinvR(c::B200Cache; scale=1.0) = _schur(c, RTLINVR, scale)             # :9-30
create_CLinvCT(c::B200Cache; scale=1.0) = _schur(c, CLINVCT, scale)             # :40-61
create_GLinvD(c::B200Cache; scale=1.0) = _schur(c, GLINVD, scale)               # :135-155
create_GLinvD_cross(c::B200Cache; scale=1.0) = _schur(c, GLINVD_CROSS, scale)   # :195-215
function create_nRTRn(c::B200Cache{N}; scale=1.0) where {N}                     # :225-244
    A = Matrix{Float64}(undef, N, N)
    check(ccall((:ilm_create_nRTRn, lib), Cint, (Ptr{Cvoid}, Cdouble, PD), c.plan, scale, A)); A
end
function create_surface_filter(c::B200Cache{N}) where {N}                       # :254-268
    A = Matrix{Float64}(undef, N, N)
    check(ccall((:ilm_create_surface_filter, lib), Cint, (Ptr{Cvoid}, PD), c.plan, A)); A
end

# ---- integrating factor / other kernels on the same engine (plan_intfact via ConstrainedSystems)
function intfact_kernel(c::B200Cache, table::Matrix{Float64}; c0=0.0, factor=1.0)
    id = Ref{Cint}(0)
    check(ccall((:ilm_add_kernel, lib), Cint, (Ptr{Cvoid}, PD, Cint, Cdouble, Cdouble, Ref{Cint}), c.plan, table, size(table, 1), c0, factor, id))
    return id[]
end
convolve!(w::GridData, c::B200Cache, kernel_id::Integer) =
    (check(ccall((:ilm_convolve, lib), Cint, (Ptr{Cvoid}, Cint, Cint, PD), c.plan, kernel_id, layout(w), w.data)); w)
function create_RTHR(c::B200Cache{N}, kernel_id::Integer; scale=1.0) where {N}   # S_i = -E H_i R of an IF-HERK stage
    A = Matrix{Float64}(undef, N, N)
    check(ccall((:ilm_create_schur_kernel, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint, PD), c.plan, RTLINVR, kernel_id, scale, 0, N, A)); A
end

# ---- surface-point solve: S \ s on the GPU (test/literate/dirichlet.jl:99)
function lu_b200(S::Matrix{Float64})
    n = size(S, 1); LU = copy(S); ipiv = Vector{Cint}(undef, n)
    check(ccall((:ilm_dense_factor, lib), Cint, (Cint, PD, Ptr{Cint}, Ptr{Cvoid}), n, LU, ipiv, C_NULL))
    return (LU, ipiv)
end
solve_b200!(b::Vector{Float64}, (LU, ipiv)) =
    (check(ccall((:ilm_dense_solve, lib), Cint, (Cint, PD, Ptr{Cint}, Cint, PD, Ptr{Cvoid}), size(LU, 1), LU, ipiv, 1, b, C_NULL)); b)

# ---- convective terms (src/grid_operators.jl:258-434); the extra caches become unused
convective_derivative!(udp::Nodes{Primal}, u::Edges{Primal}, p::Nodes{Primal}, c::B200Cache, extra=nothing) =
    (check(ccall((:ilm_convective_derivative_scalar, lib), Cint, (Ptr{Cvoid}, PD, PD, PD), c.plan, u.data, p.data, udp.data)); udp)
convective_derivative!(vdw::Nodes{Dual}, v::Edges{Primal}, w::Nodes{Dual}, c::B200Cache, extra=nothing) =
    (check(ccall((:ilm_convective_derivative_dual, lib), Cint, (Ptr{Cvoid}, PD, PD, PD), c.plan, v.data, w.data, vdw.data)); vdw)
convective_derivative!(vdu::Edges{Primal}, v::Edges{Primal}, u::Edges{Primal}, c::B200Cache, extra=nothing) =
    (check(ccall((:ilm_convective_derivative_vector, lib), Cint, (Ptr{Cvoid}, PD, PD, PD), c.plan, v.data, u.data, vdu.data)); vdu)
convective_derivative!(udu::Edges{Primal}, u::Edges{Primal}, c::B200Cache, extra=nothing) =
    convective_derivative!(udu, u, u, c, extra)
w_cross_v!(vw::Edges{Primal}, w::Nodes{Dual}, v::Edges{Primal}, c::B200Cache, extra=nothing) =
    (check(ccall((:ilm_w_cross_v, lib), Cint, (Ptr{Cvoid}, PD, PD, PD), c.plan, w.data, v.data, vw.data)); vw)

# ---- forcing regions (src/forcing.jl): region caches share the base cache's Laplacian (`L = L`, :201-248)
"""
    B200Cache(region::BasicILMCache, parent::B200Cache; ddftype)

Device plan for a region cache built with the parent's `L` (`AreaRegionCache`, `LineRegionCache`): aliases the
parent's multipliers and scratch (ilm_plan_create_shared); the child holds a reference to the parent, so the
parent's finalizer cannot run first.
"""
function B200Cache(cache::BasicILMCache{N,SCA}, parent::B200Cache; ddftype=CartesianGrids.Yang3) where {N,SCA}
    pts, nrm, ds = points(cache), normals(cache), areas(cache)
    plan = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ilm_plan_create_shared, lib), Cint,
        (Ptr{Cvoid}, Cint, PD, PD, PD, PD, PD, Cint, Cint, Ref{Ptr{Cvoid}}),
        parent.plan, N, pts.u, pts.v, nrm.u, nrm.v, ds.data, DDF[ddftype], SCA == GridScaling ? 0 : 1, plan))
    c = B200Cache{N,SCA,typeof(cache)}(cache, plan[], cache.gdata_cache isa Edges, parent)   # keeps the parent plan alive
    finalizer(x -> ccall((:ilm_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), x.plan), c)
    return c
end
# _apply_forcing! bodies (src/forcing.jl:456-494): dy .+= str .* mask   and   dy .+= R str
forcing_area_add!(dy::GridData, str::GridData, msk::Union{GridData,Nothing}, c::B200Cache) =
    (check(ccall((:ilm_forcing_area_add, lib), Cint, (Ptr{Cvoid}, Cint, PD, PD, PD), c.plan, layout(dy), str.data,
                 msk === nothing ? C_NULL : msk.data, dy.data)); dy)
forcing_line_add!(dy::GridData, str::PointData, c::B200Cache) =
    (check(ccall((:ilm_forcing_line_add, lib), Cint, (Ptr{Cvoid}, Cint, PD, PD), c.plan, layout(dy), str.data, dy.data)); dy)

# ---- Helmholtz decomposition (src/helmholtz.jl); the extra caches (wcache, dcache, veccache) become unused
import ImmersedLayers: vectorpotential_from_masked_curlv!, scalarpotential_from_masked_divv!, vectorpotential_from_curlv!,
    scalarpotential_from_divv!, vecfield_from_vectorpotential!, vecfield_from_scalarpotential!, vecfield_helmholtz!,
    masked_curlv_from_curlv_masked!, curlv_masked_from_masked_curlv!, masked_divv_from_divv_masked!, divv_masked_from_masked_divv!
_pot(c, curlv, divv, dv, psi, phi) = check(ccall((:ilm_helmholtz_potentials, lib), Cint, (Ptr{Cvoid}, PD, PD, PD, PD, PD),
                                                 c.plan, curlv, divv, dv, psi, phi))
vectorpotential_from_masked_curlv!(ψ::Nodes{Dual}, curlv::Nodes{Dual}, dv::VectorData, c::B200Cache, wcache=nothing) =
    (_pot(c, curlv.data, C_NULL, dv.data, ψ.data, C_NULL); ψ)                                   # :84-96
scalarpotential_from_masked_divv!(ϕ::Nodes{Primal}, divv::Nodes{Primal}, dv::VectorData, c::B200Cache, dcache=nothing) =
    (_pot(c, C_NULL, divv.data, dv.data, C_NULL, ϕ.data); ϕ)                                    # :186-201
vectorpotential_from_curlv!(ψ::Nodes{Dual}, curlv::Nodes{Dual}, c::B200Cache) = (_pot(c, curlv.data, C_NULL, C_NULL, ψ.data, C_NULL); ψ)
scalarpotential_from_divv!(ϕ::Nodes{Primal}, divv::Nodes{Primal}, c::B200Cache) = (_pot(c, C_NULL, divv.data, C_NULL, C_NULL, ϕ.data); ϕ)
_vfp(c, psi, phi, vp, v) = check(ccall((:ilm_vecfield_from_potentials, lib), Cint, (Ptr{Cvoid}, PD, PD, PD, PD), c.plan, psi, phi, vp, v))
vecfield_from_vectorpotential!(v::Edges{Primal}, ψ::Nodes{Dual}, c::B200Cache) = (_vfp(c, ψ.data, C_NULL, C_NULL, v.data); v)
vecfield_from_scalarpotential!(v::Edges{Primal}, ϕ::Nodes{Primal}, c::B200Cache) = (_vfp(c, C_NULL, ϕ.data, C_NULL, v.data); v)
vecfield_helmholtz!(v::Edges{Primal}, curlv::Nodes{Dual}, divv::Nodes{Primal}, dv::VectorData, vp::Union{Edges{Primal},Nothing},
                    c::B200Cache, veccache=nothing) =                                             # :285-307
    (check(ccall((:ilm_vecfield_helmholtz, lib), Cint, (Ptr{Cvoid}, PD, PD, PD, PD, PD), c.plan, curlv.data, divv.data, dv.data,
                 vp === nothing ? C_NULL : vp.data, v.data)); v)
_jump(c, op, sign, dv, fin, fout) = check(ccall((:ilm_helmholtz_jump_add, lib), Cint, (Ptr{Cvoid}, Cint, Cint, PD, PD, PD),
                                                c.plan, op, sign, dv.data, fin.data, fout.data))
masked_curlv_from_curlv_masked!(mw::Nodes{Dual}, w::Nodes{Dual}, dv::VectorData, c::B200Cache, wcache=nothing) = (_jump(c, 0, -1, dv, w, mw); mw)
curlv_masked_from_masked_curlv!(w::Nodes{Dual}, mw::Nodes{Dual}, dv::VectorData, c::B200Cache, wcache=nothing) = (_jump(c, 0, 1, dv, mw, w); w)
masked_divv_from_divv_masked!(md::Nodes{Primal}, d::Nodes{Primal}, dv::VectorData, c::B200Cache, dcache=nothing) = (_jump(c, 1, -1, dv, d, md); md)
divv_masked_from_masked_divv!(d::Nodes{Primal}, md::Nodes{Primal}, dv::VectorData, c::B200Cache, dcache=nothing) = (_jump(c, 1, 1, dv, md, d); d)

# ---- whole-problem entry point and multi-GPU (include/ilm_b200.h: ilm_dirichlet_poisson, ilm_comm_*, ilm_create_schur_sharded)
"""
    dirichlet_solve!(f::Nodes{Primal}, s::ScalarData, fplus::ScalarData, fminus, c::B200Cache)

`solve(prob::DirichletPoissonProblem, sys)` of test/literate/dirichlet.jl:71-107 in one library call; S and its LU
factors stay on the device.  With a communicator on the plan (`comm_init!`) the Schur columns are sharded.
"""
dirichlet_solve!(f::Union{Nodes{Primal},Nothing}, s::ScalarData, fplus::ScalarData, fminus::Union{ScalarData,Nothing}, c::B200Cache) =
    (check(ccall((:ilm_dirichlet_poisson, lib), Cint, (Ptr{Cvoid}, PD, PD, PD, PD, PD), c.plan, fplus.data,
                 fminus === nothing ? C_NULL : fminus.data, f === nothing ? C_NULL : f.data, s.data, C_NULL)); (f, s))   # f = nothing: multiplier only (sharded solve, field on one rank)
# rows [row0, row1) (0-based, half-open) of the field only: every rank of a sharded solve keeps a slab of the result
dirichlet_solve_rows!(frows::AbstractVector{Float64}, row0::Integer, row1::Integer, s::ScalarData, fplus::ScalarData,
                      fminus::Union{ScalarData,Nothing}, c::B200Cache) =
    (check(ccall((:ilm_dirichlet_poisson_rows, lib), Cint, (Ptr{Cvoid}, PD, PD, PD, Cint, Cint, PD, PD), c.plan, fplus.data,
                 fminus === nothing ? C_NULL : fminus.data, frows, row0, row1, s.data, C_NULL)); (frows, s))
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:ilm_comm_unique_id, lib), Cint, (Ptr{UInt8}, Cint), id, 128)); id)
# `id` from rank 0, distributed by the host program (MPI.Bcast!, a file, ...)
comm_init!(c::B200Cache, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    (check(ccall((:ilm_comm_init, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint), c.plan, id, 128, rank, nranks)); c)
comm_destroy!(c::B200Cache) = (check(ccall((:ilm_comm_destroy, lib), Cint, (Ptr{Cvoid},), c.plan)); c)
function create_schur_sharded(c::B200Cache{N}, which::Integer; scale=1.0, kernel_id=0) where {N}
    A = Matrix{Float64}(undef, N, N)
    check(ccall((:ilm_create_schur_sharded, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, PD), c.plan, which, kernel_id, scale, A)); A
end

end # module

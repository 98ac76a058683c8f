// ilm_api_vector.cu -- C ABI of the vector-data cache (SurfaceVectorCache,
// src/cache.jl:184-202): TensorData / EdgeGradient operators, the vector forms
// of the surface-grid composites and the 2N x 2N Schur builders
// (src/surface_operators.jl:113-138,172-215,242-343,388-398,443-452,559-664;
// src/matrix_operators.jl:9-30,40-61,102-185,225-244).
#include <algorithm>

#include "ilm_internal.h"

using namespace ilm;

namespace {

struct Sizes {
    size_t Pn, Pd, nu, nv, ne, ng;
    explicit Sizes(const ilm_plan* p) {
        const int NX = p->g.NX, NY = p->g.NY;
        Pn = (size_t)(NX - 1) * (NY - 1); Pd = (size_t)NX * NY;
        nu = (size_t)NX * (NY - 1); nv = (size_t)(NX - 1) * NY;
        ne = nu + nv; ng = 2 * Pn + 2 * Pd;
    }
};

double deriv_div(const ilm_plan* p) { return p->scaling == ILM_GRID_SCALING ? p->g.dx : 1.0; }

int need_tensor_scratch(ilm_plan* p) {
    if (!p->g_tensor) ILM_CUDA(cudaMalloc(&p->g_tensor, Sizes(p).ng * sizeof(double)));
    return ILM_OK;
}

// EdgeGradient <- TensorData (4 tables: primal, dual, dual, primal)
int regularize_tensor(ilm_plan* p, const double* T, double* eg) {
    const Sizes z(p);
    const int N = p->N;
    ILM_TRY(launch_regularize(p, p->tab[ILM_NODES_PRIMAL], T, nullptr, 1.0, eg, true));
    ILM_TRY(launch_regularize(p, p->tab[ILM_NODES_DUAL], T + N, nullptr, 1.0, eg + z.Pn, true));
    ILM_TRY(launch_regularize(p, p->tab[ILM_NODES_DUAL], T + 2 * (size_t)N, nullptr, 1.0, eg + z.Pn + z.Pd, true));
    ILM_TRY(launch_regularize(p, p->tab[ILM_NODES_PRIMAL], T + 3 * (size_t)N, nullptr, 1.0, eg + z.Pn + 2 * z.Pd, true));
    return ILM_OK;
}
int interpolate_tensor(ilm_plan* p, const double* eg, double* T) {
    const Sizes z(p);
    const int N = p->N;
    ILM_TRY(launch_interpolate(p, p->tab[ILM_NODES_PRIMAL], eg, T));
    ILM_TRY(launch_interpolate(p, p->tab[ILM_NODES_DUAL], eg + z.Pn, T + N));
    ILM_TRY(launch_interpolate(p, p->tab[ILM_NODES_DUAL], eg + z.Pn + z.Pd, T + 2 * (size_t)N));
    ILM_TRY(launch_interpolate(p, p->tab[ILM_NODES_PRIMAL], eg + z.Pn + 2 * z.Pd, T + 3 * (size_t)N));
    return ILM_OK;
}
int regularize_edges(ilm_plan* p, const double* v, double* edges) {
    ILM_TRY(launch_regularize(p, p->tab[ILM_XEDGES], v, nullptr, 1.0, edges, true));
    return launch_regularize(p, p->tab[ILM_YEDGES], v + p->N, nullptr, 1.0, edges + Sizes(p).nu, true);
}
int interpolate_edges(ilm_plan* p, const double* edges, double* v) {
    ILM_TRY(launch_interpolate(p, p->tab[ILM_XEDGES], edges, v));
    return launch_interpolate(p, p->tab[ILM_YEDGES], edges + Sizes(p).nu, v + p->N);
}

// device composites (inputs/outputs are device pointers)
int regularize_normal_tensor_dev(ilm_plan* p, int mode, const double* v, double* eg) {
    ILM_TRY(launch_tensor_from_vector(p, mode, v, p->s_b));
    return regularize_tensor(p, p->s_b, eg);
}
int normal_interpolate_tensor_dev(ilm_plan* p, int mode, const double* eg, double* v, double div) {
    ILM_TRY(interpolate_tensor(p, eg, p->s_b));
    return launch_tensor_dot(p, mode, p->s_b, div, v);
}
int vsurface_divergence_dev(ilm_plan* p, int mode, const double* v, double* edges) {
    ILM_TRY(need_tensor_scratch(p));
    ILM_TRY(regularize_normal_tensor_dev(p, mode, v, p->g_tensor));
    return launch_div_tensor(p, p->g_tensor, edges, deriv_div(p));
}
int vsurface_grad_dev(ilm_plan* p, int mode, const double* edges, double* v) {
    ILM_TRY(need_tensor_scratch(p));
    ILM_TRY(launch_grad_tensor(p, edges, p->g_tensor, 1.0));
    return normal_interpolate_tensor_dev(p, mode, p->g_tensor, v, deriv_div(p));
}
int vsurface_curl_s2n_dev(ilm_plan* p, const double* v, double* dual) {
    ILM_TRY(regularize_edges(p, v, p->g_edges));
    return launch_curl_e2n(p, p->g_edges, p->g_edges + Sizes(p).nu, dual, deriv_div(p));
}
__global__ void k_div_scalar(double* __restrict__ w, int n, double div) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = w[i] / div;
}
int vsurface_curl_n2s_dev(ilm_plan* p, const double* dual, double* v) {
    ILM_TRY(launch_curl_n2e(p, dual, p->g_edges, p->g_edges + Sizes(p).nu, 1.0));
    ILM_TRY(interpolate_edges(p, p->g_edges, v));
    if (p->N && deriv_div(p) != 1.0) {
        k_div_scalar<<<(2 * p->N + 127) / 128, 128, 0, p->stream>>>(v, 2 * p->N, deriv_div(p));
        ILM_CUDA(cudaGetLastError());
        p->launches++;
    }
    return ILM_OK;
}

__global__ void k_unit2(double* __restrict__ s, int n, int col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = (i == col) ? 1.0 : 0.0;
}

}  // namespace

#define ILM_VPLAN(p)                                           \
    do {                                                       \
        if (!(p)) { set_error("null plan"); return ILM_EINVAL; } \
        cudaSetDevice((p)->device);                            \
    } while (0)
#define ILM_TMODE(m) \
    if ((m) != ILM_TENSOR_NORMAL && (m) != ILM_TENSOR_SYMM) { set_error("bad tensor mode"); return ILM_EINVAL; }

extern "C" int ilm_regularize_normal_tensor(ilm_plan* p, int mode, const double* v, double* eg) {
    ILM_VPLAN(p);
    ILM_TMODE(mode);
    Io io(p);
    const double* dv = io.in(v, 2 * (size_t)p->N);
    double* de = io.out(eg, Sizes(p).ng);
    if (io.status) return io.status;
    ILM_TRY(regularize_normal_tensor_dev(p, mode, dv, de));
    return io.finish();
}

extern "C" int ilm_normal_interpolate_tensor(ilm_plan* p, int mode, const double* eg, double* v) {
    ILM_VPLAN(p);
    ILM_TMODE(mode);
    Io io(p);
    const double* de = io.in(eg, Sizes(p).ng);
    double* dv = io.out(v, 2 * (size_t)p->N);
    if (io.status) return io.status;
    ILM_TRY(normal_interpolate_tensor_dev(p, mode, de, dv, 1.0));
    return io.finish();
}

extern "C" int ilm_regularize_normal_vs(ilm_plan* p, int op, const double* v, double* nodes) {
    ILM_VPLAN(p);
    if (op != ILM_VS_CROSS && op != ILM_VS_DOT) { set_error("bad op"); return ILM_EINVAL; }
    const Sizes z(p);
    Io io(p);
    const double* dv = io.in(v, 2 * (size_t)p->N);
    double* dn = io.out(nodes, op == ILM_VS_CROSS ? z.Pd : z.Pn);
    if (io.status) return io.status;
    ILM_TRY(launch_vec_pointwise(p, op == ILM_VS_CROSS ? 0 : 1, dv, p->s_b));
    ILM_TRY(launch_regularize(p, p->tab[op == ILM_VS_CROSS ? ILM_NODES_DUAL : ILM_NODES_PRIMAL], p->s_b, nullptr, 1.0, dn, true));
    return io.finish();
}

extern "C" int ilm_normal_interpolate_vs(ilm_plan* p, int op, const double* nodes, double* v) {
    ILM_VPLAN(p);
    if (op != ILM_VS_CROSS && op != ILM_VS_DOT) { set_error("bad op"); return ILM_EINVAL; }
    const Sizes z(p);
    Io io(p);
    const double* dn = io.in(nodes, op == ILM_VS_CROSS ? z.Pd : z.Pn);
    double* dv = io.out(v, 2 * (size_t)p->N);
    if (io.status) return io.status;
    ILM_TRY(launch_interpolate(p, p->tab[op == ILM_VS_CROSS ? ILM_NODES_DUAL : ILM_NODES_PRIMAL], dn, p->s_b));
    ILM_TRY(launch_vec_pointwise(p, op == ILM_VS_CROSS ? 2 : 3, p->s_b, dv));
    return io.finish();
}

extern "C" int ilm_regularize_normal_dot_tensor(ilm_plan* p, const double* tau, double* edges) {
    ILM_VPLAN(p);
    Io io(p);
    const double* dt = io.in(tau, 4 * (size_t)p->N);
    double* de = io.out(edges, Sizes(p).ne);
    if (io.status) return io.status;
    ILM_TRY(launch_tensor_dot(p, ILM_TENSOR_NORMAL, dt, 1.0, p->s_b));
    ILM_TRY(regularize_edges(p, p->s_b, de));
    return io.finish();
}

extern "C" int ilm_normal_dot_interpolate_tensor(ilm_plan* p, const double* edges, double* tau) {
    ILM_VPLAN(p);
    Io io(p);
    const double* de = io.in(edges, Sizes(p).ne);
    double* dt = io.out(tau, 4 * (size_t)p->N);
    if (io.status) return io.status;
    ILM_TRY(interpolate_edges(p, de, p->s_b));
    ILM_TRY(launch_tensor_from_vector(p, ILM_TENSOR_NORMAL, p->s_b, dt));
    return io.finish();
}

extern "C" int ilm_grad_tensor(ilm_plan* p, const double* edges, double* eg) {
    ILM_VPLAN(p);
    Io io(p);
    const double* de = io.in(edges, Sizes(p).ne);
    double* dg = io.out(eg, Sizes(p).ng);
    if (io.status) return io.status;
    ILM_TRY(launch_grad_tensor(p, de, dg, deriv_div(p)));
    return io.finish();
}

extern "C" int ilm_divergence_tensor(ilm_plan* p, const double* eg, double* edges) {
    ILM_VPLAN(p);
    Io io(p);
    const double* dg = io.in(eg, Sizes(p).ng);
    double* de = io.out(edges, Sizes(p).ne);
    if (io.status) return io.status;
    ILM_TRY(launch_div_tensor(p, dg, de, deriv_div(p)));
    return io.finish();
}

extern "C" int ilm_vsurface_divergence(ilm_plan* p, int mode, const double* v, double* edges) {
    ILM_VPLAN(p);
    ILM_TMODE(mode);
    Io io(p);
    const double* dv = io.in(v, 2 * (size_t)p->N);
    double* de = io.out(edges, Sizes(p).ne);
    if (io.status) return io.status;
    ILM_TRY(vsurface_divergence_dev(p, mode, dv, de));
    return io.finish();
}

extern "C" int ilm_vsurface_grad(ilm_plan* p, int mode, const double* edges, double* v) {
    ILM_VPLAN(p);
    ILM_TMODE(mode);
    Io io(p);
    const double* de = io.in(edges, Sizes(p).ne);
    double* dv = io.out(v, 2 * (size_t)p->N);
    if (io.status) return io.status;
    ILM_TRY(vsurface_grad_dev(p, mode, de, dv));
    return io.finish();
}

extern "C" int ilm_vsurface_curl_s2n(ilm_plan* p, const double* v, double* dual) {
    ILM_VPLAN(p);
    Io io(p);
    const double* dv = io.in(v, 2 * (size_t)p->N);
    double* dn = io.out(dual, Sizes(p).Pd);
    if (io.status) return io.status;
    ILM_TRY(vsurface_curl_s2n_dev(p, dv, dn));
    return io.finish();
}

extern "C" int ilm_vsurface_curl_n2s(ilm_plan* p, const double* dual, double* v) {
    ILM_VPLAN(p);
    Io io(p);
    const double* dn = io.in(dual, Sizes(p).Pd);
    double* dv = io.out(v, 2 * (size_t)p->N);
    if (io.status) return io.status;
    ILM_TRY(vsurface_curl_n2s_dev(p, dn, dv));
    return io.finish();
}

namespace ilm {
int mask_edges_dev(ilm_plan* p, double* de) {
    const Sizes z(p);
    if (p->N == 0) return launch_fill(p, de, z.ne, 1.0);
    ILM_TRY(launch_fill(p, p->s_a, 2 * (size_t)p->N, 1.0));
    ILM_TRY(vsurface_divergence_dev(p, ILM_TENSOR_NORMAL, p->s_a, de));
    ILM_TRY(conv_apply(p, 0, FieldRef{de, p->g.NX, p->g.NY - 1}, FieldRef{de + z.nu, p->g.NX - 1, p->g.NY}));
    return launch_scale(p, de, z.ne, -1.0);
}
}  // namespace ilm

extern "C" int ilm_mask_edges(ilm_plan* p, double* edges) {
    ILM_VPLAN(p);
    Io io(p);
    double* de = io.out(edges, Sizes(p).ne);
    if (io.status) return io.status;
    ILM_TRY(mask_edges_dev(p, de));
    return io.finish();
}

// mask!(w, cache) / complementary_mask!(w, cache) (src/surface_operators.jl:788-823, 880-923): the
// mask is recomputed (one inverse Laplacian, as the reference does) and multiplied into w, averaged
// onto w's layout where that differs from the cache's own grid-data layout.
extern "C" int ilm_mask_product(ilm_plan* p, int cache_kind, int layout, int complementary, double* w) {
    ILM_VPLAN(p);
    if (p->scaling != ILM_GRID_SCALING) { set_error("mask!: only defined for GridScaling caches"); return ILM_EINVAL; }
    if (cache_kind != ILM_SCALAR_CACHE && cache_kind != ILM_VECTOR_CACHE) { set_error("mask!: bad cache kind"); return ILM_EINVAL; }
    const bool vec = cache_kind == ILM_VECTOR_CACHE;
    const bool ok = layout >= ILM_NODES_PRIMAL && (vec ? (layout == ILM_NODES_PRIMAL || layout == ILM_NODES_DUAL || layout == ILM_EDGES || layout == ILM_EDGEGRAD)
                                                       : layout <= ILM_EDGES);
    if (!ok) { set_error("mask!: no method for this grid data type on this cache"); return ILM_EINVAL; }
    const Sizes z(p);
    const size_t n = layout == ILM_EDGES ? z.ne : layout == ILM_EDGEGRAD ? z.ng : layout_info(layout, p->g.NX, p->g.NY).n();
    Io io(p);
    double* dw = io.inout(w, n);
    if (io.status) return io.status;
    const int c = complementary ? 1 : 0;
    if (!vec) {
        double* m = p->g_b;                                   // Nodes{Primal} mask (gdata_cache)
        ILM_TRY(mask_primal_dev(p, m));
        if (layout == ILM_EDGES) {
            ILM_TRY(launch_mask_product(p, dw, ILM_XEDGES, m, ILM_NODES_PRIMAL, c));
            ILM_TRY(launch_mask_product(p, dw + z.nu, ILM_YEDGES, m, ILM_NODES_PRIMAL, c));
        } else {
            ILM_TRY(launch_mask_product(p, dw, layout, m, ILM_NODES_PRIMAL, c));
        }
    } else {
        double* m = p->g_edges;                               // Edges mask (gdata_cache of the vector cache)
        ILM_TRY(mask_edges_dev(p, m));
        if (layout == ILM_EDGES) {                            // product!(w, gdata_cache, w)
            ILM_TRY(launch_mask_product(p, dw, ILM_XEDGES, m, ILM_XEDGES, c));
            ILM_TRY(launch_mask_product(p, dw + z.nu, ILM_YEDGES, m + z.nu, ILM_YEDGES, c));
        } else if (layout == ILM_EDGEGRAD) {                  // every component uses gdata_cache.u (:908-923)
            ILM_TRY(launch_mask_product(p, dw, ILM_NODES_PRIMAL, m, ILM_XEDGES, c));
            ILM_TRY(launch_mask_product(p, dw + z.Pn, ILM_NODES_DUAL, m, ILM_XEDGES, c));
            ILM_TRY(launch_mask_product(p, dw + z.Pn + z.Pd, ILM_NODES_DUAL, m, ILM_XEDGES, c));
            ILM_TRY(launch_mask_product(p, dw + z.Pn + 2 * z.Pd, ILM_NODES_PRIMAL, m, ILM_XEDGES, c));
        } else {
            ILM_TRY(launch_mask_product(p, dw, layout, m, ILM_XEDGES, c));
        }
    }
    return io.finish();
}

extern "C" int ilm_create_schur_vector(ilm_plan* p, int which, double scale, int col_begin, int col_end, double* A) {
    ILM_VPLAN(p);
    const int N = p->N, M = 2 * N;
    if (which < ILM_V_RTLINVR || which > ILM_V_GLINVD_SYMM) { set_error("ilm_create_schur_vector: unknown matrix"); return ILM_EINVAL; }
    if (col_begin < 0 || col_end > M || col_begin > col_end) { set_error("ilm_create_schur_vector: bad column range"); return ILM_ESIZE; }
    const int ncols = col_end - col_begin;
    if (N == 0 || ncols == 0) return ILM_OK;
    const Sizes z(p);
    Io io(p);
    double* dA = io.out(A, (size_t)M * ncols);
    if (io.status) return io.status;
    double* unit = p->s_a;            // 2N
    double* sout = p->s_a + M;        // 2N
    const int mode = which == ILM_V_GLINVD_SYMM ? ILM_TENSOR_SYMM : ILM_TENSOR_NORMAL;
    const bool curl = which == ILM_V_CLINVCT || which == ILM_V_CL2INVCT;
    // grid scratch: Nodes{Dual} right-hand sides (curl builders) live in g_a; Edges-valued ones in g_edges,
    // whose u and v components ride one complex transform
    double* fu = p->g_a;
    if (curl) {
        // Curl builders: the right-hand side C_s^T e_c is a Nodes{Dual} field, so TWO columns ride one complex
        // transform (as in the scalar builders).  R_f e_c is a W x W edge patch, its curl reaches one row further:
        // the first transform gets that row range as sparse input; the last one only inverts the rows that
        // C_s (curl + edge interpolation) can read.
        const DevTable& tu = p->tab[ILM_XEDGES];
        const DevTable& tv = p->tab[ILM_YEDGES];
        auto rows_of = [&](int c, int* lo, int* hi) {
            const DevTable& t = c < N ? tu : tv;
            const int k = c < N ? c : c - N;
            *lo = std::min(*lo, t.h_j0[k] - 2);
            *hi = std::max(*hi, t.h_j0[k] + t.W + 2);
        };
        int olo, ohi;
        probe_output_rows(p, ILM_CLINVCT, &olo, &ohi);
        const int reps = which == ILM_V_CL2INVCT ? 2 : 1;
        double* fv = p->g_b;
        for (int c = col_begin; c < col_end; c += 2) {
            const bool two = c + 1 < col_end;
            int rlo = 1 << 30, rhi = -1;
            for (int q = 0; q < (two ? 2 : 1); ++q) rows_of(c + q, &rlo, &rhi);
            rlo = std::max(rlo, 0); rhi = std::min(rhi, p->g.NY);
            if (rhi <= rlo) { rlo = 0; rhi = 1; }          // every window of the pair lies off the grid: an empty (zero) patch row
            ILM_TRY(launch_vcurl_probe_pre(p, c, two ? 2 : 1, fu, fv, rlo, rhi, deriv_div(p)));   // C^T R_f e_c / dx on the patch rows
            const FieldRef f1{fu, p->g.NX, p->g.NY};
            const FieldRef f2 = two ? FieldRef{fv, p->g.NX, p->g.NY} : FieldRef{nullptr, 0, 0};
            for (int r = 0; r < reps; ++r) {
                const bool first = r == 0, last = r == reps - 1;
                ILM_TRY(conv_apply(p, 0, f1, f2, first ? rlo : -1, first ? rhi : -1, last ? olo : -1, last ? ohi : -1));
            }
            ILM_TRY(launch_vcurl_probe_post(p, two ? 2 : 1, fu, fv, deriv_div(p), -scale, dA + (size_t)(c - col_begin) * M,
                                            dA + (size_t)(c + 1 - col_begin) * M));
        }
        return io.finish();
    }
    for (int c = col_begin; c < col_end; ++c) {
        k_unit2<<<(M + 127) / 128, 128, 0, p->stream>>>(unit, M, c);
        ILM_CUDA(cudaGetLastError());
        p->launches++;
        {
            // Edges-valued field: stage through g_edges, then split into (fu, fv) for the pair transform
            if (which == ILM_V_RTLINVR) ILM_TRY(regularize_edges(p, unit, p->g_edges));
            else ILM_TRY(vsurface_divergence_dev(p, mode, unit, p->g_edges));
            ILM_TRY(conv_apply(p, 0, FieldRef{p->g_edges, p->g.NX, p->g.NY - 1},
                               FieldRef{p->g_edges + z.nu, p->g.NX - 1, p->g.NY}));
            if (which == ILM_V_RTLINVR) {
                ILM_TRY(interpolate_edges(p, p->g_edges, sout));
            } else {
                // surface_grad reads Edges and uses g_tensor as scratch; g_edges is the input here
                ILM_TRY(need_tensor_scratch(p));
                ILM_TRY(launch_grad_tensor(p, p->g_edges, p->g_tensor, 1.0));
                ILM_TRY(normal_interpolate_tensor_dev(p, mode, p->g_tensor, sout, deriv_div(p)));
            }
        }
        ILM_TRY(launch_scale_store_column(p, sout, dA + (size_t)(c - col_begin) * M, M, -scale));
    }
    return io.finish();
}

extern "C" int ilm_create_nRTRn_vector(ilm_plan* p, double scale, double* A) {
    ILM_VPLAN(p);
    const int N = p->N, M = 2 * N;
    if (N == 0) return ILM_OK;
    Io io(p);
    double* dA = io.out(A, (size_t)M * M);
    if (io.status) return io.status;
    ILM_TRY(need_tensor_scratch(p));
    double* unit = p->s_a;
    double* sout = p->s_a + M;
    for (int c = 0; c < M; ++c) {
        k_unit2<<<(M + 127) / 128, 128, 0, p->stream>>>(unit, M, c);
        ILM_CUDA(cudaGetLastError());
        p->launches++;
        ILM_TRY(regularize_normal_tensor_dev(p, ILM_TENSOR_NORMAL, unit, p->g_tensor));
        ILM_TRY(normal_interpolate_tensor_dev(p, ILM_TENSOR_NORMAL, p->g_tensor, sout, 1.0));
        ILM_TRY(launch_scale_store_column(p, sout, dA + (size_t)c * M, M, scale));
    }
    return io.finish();
}

// ---------------------------------------------------------------- Helmholtz decomposition (src/helmholtz.jl)
namespace {
// out = in + sign * R (n x dv) on Nodes{Dual} (op = ILM_VS_CROSS) or in + sign * R (n . dv) on Nodes{Primal}
// (op = ILM_VS_DOT); `in` may alias `out` or be null (zero field).  The reference zero-fills `out`, regularizes,
// scales by -1 where needed and adds `in` as whole-field sweeps (:84-96, 149-160, 171-181, 186-201, 240-270).
int jump_add_dev(ilm_plan* p, int op, bool negate, const double* dv, const double* in, double* out) {
    const Sizes z(p);
    const size_t n = op == ILM_VS_CROSS ? z.Pd : z.Pn;
    if (!in) ILM_TRY(launch_fill(p, out, n, 0.0));
    else if (in != out) ILM_CUDA(cudaMemcpyAsync(out, in, n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    if (p->N == 0 || !dv) return ILM_OK;           // VectorData{0} methods: the field is passed through
    ILM_TRY(launch_vec_pointwise(p, op == ILM_VS_CROSS ? 0 : 1, dv, p->s_b));
    return launch_regularize_add(p, p->tab[op == ILM_VS_CROSS ? ILM_NODES_DUAL : ILM_NODES_PRIMAL], p->s_b, out, negate);
}

// psi = -L^-1 (masked_curlv + R n x [v]) (:84-96, 121-124), phi = L^-1 (masked_divv + R n . [v]) (:186-201, 216-219);
// both right-hand sides ride ONE complex transform.  Either output may be null.
int helmholtz_potentials_dev(ilm_plan* p, const double* curlv, const double* divv, const double* dv, double* psi, double* phi) {
    const Sizes z(p);
    const int NX = p->g.NX, NY = p->g.NY;
    if (psi) ILM_TRY(jump_add_dev(p, ILM_VS_CROSS, false, dv, curlv, psi));
    if (phi) ILM_TRY(jump_add_dev(p, ILM_VS_DOT, false, dv, divv, phi));
    const FieldRef fpsi{psi, NX, NY}, fphi{phi, NX - 1, NY - 1}, none{nullptr, 0, 0};
    if (psi && phi) ILM_TRY(conv_apply(p, 0, fpsi, fphi));
    else if (psi) ILM_TRY(conv_apply(p, 0, fpsi, none));
    else if (phi) ILM_TRY(conv_apply(p, 0, fphi, none));
    if (psi) ILM_TRY(launch_scale(p, psi, z.Pd, -1.0));
    return ILM_OK;
}
}  // namespace

/* masked_curlv_from_curlv_masked! (sign -1), curlv_masked_from_masked_curlv! (+1) on Nodes{Dual} with op = ILM_VS_CROSS;
 * masked_divv_from_divv_masked! (-1), divv_masked_from_masked_divv! (+1) on Nodes{Primal} with op = ILM_VS_DOT */
extern "C" int ilm_helmholtz_jump_add(ilm_plan* p, int op, int sign, const double* dv, const double* in, double* out) {
    ILM_VPLAN(p);
    if ((op != ILM_VS_CROSS && op != ILM_VS_DOT) || (sign != 1 && sign != -1) || !out) { set_error("ilm_helmholtz_jump_add: bad arguments"); return ILM_EINVAL; }
    const Sizes z(p);
    const size_t n = op == ILM_VS_CROSS ? z.Pd : z.Pn;
    Io io(p);
    const double* ddv = dv ? io.in(dv, 2 * (size_t)p->N) : nullptr;
    const double* din = (in && in != out) ? io.in(in, n) : nullptr;
    double* dout = (in == out) ? io.inout(out, n) : io.out(out, n);
    if (io.status) return io.status;
    ILM_TRY(jump_add_dev(p, op, sign < 0, ddv, in == out ? dout : din, dout));
    return io.finish();
}

extern "C" int ilm_helmholtz_potentials(ilm_plan* p, const double* masked_curlv, const double* masked_divv, const double* dv,
                                        double* psi, double* phi) {
    ILM_VPLAN(p);
    if (!psi && !phi) { set_error("ilm_helmholtz_potentials: no output"); return ILM_EINVAL; }
    const Sizes z(p);
    Io io(p);
    const double* dc = (psi && masked_curlv) ? io.in(masked_curlv, z.Pd) : nullptr;
    const double* dd = (phi && masked_divv) ? io.in(masked_divv, z.Pn) : nullptr;
    const double* ddv = dv ? io.in(dv, 2 * (size_t)p->N) : nullptr;
    double* dpsi = psi ? io.out(psi, z.Pd) : nullptr;
    double* dphi = phi ? io.out(phi, z.Pn) : nullptr;
    if (io.status) return io.status;
    ILM_TRY(helmholtz_potentials_dev(p, dc, dd, ddv, dpsi, dphi));
    return io.finish();
}

extern "C" int ilm_vecfield_from_potentials(ilm_plan* p, const double* psi, const double* phi, const double* vp, double* v) {
    ILM_VPLAN(p);
    if (!v) { set_error("ilm_vecfield_from_potentials: null output"); return ILM_EINVAL; }
    const Sizes z(p);
    Io io(p);
    const double* dpsi = psi ? io.in(psi, z.Pd) : nullptr;
    const double* dphi = phi ? io.in(phi, z.Pn) : nullptr;
    const double* dvp = vp ? io.in(vp, z.ne) : nullptr;
    double* dvv = io.out(v, z.ne);
    if (io.status) return io.status;
    ILM_TRY(launch_vecfield_from_potentials(p, dpsi, dphi, dvp, dvp ? dvp + z.nu : nullptr, dvv, dvv + z.nu, deriv_div(p)));
    return io.finish();
}

extern "C" int ilm_vecfield_helmholtz(ilm_plan* p, const double* masked_curlv, const double* masked_divv, const double* dv,
                                      const double* vp, double* v) {
    ILM_VPLAN(p);
    if (!v) { set_error("ilm_vecfield_helmholtz: null output"); return ILM_EINVAL; }
    const Sizes z(p);
    Io io(p);
    const double* dc = masked_curlv ? io.in(masked_curlv, z.Pd) : nullptr;
    const double* dd = masked_divv ? io.in(masked_divv, z.Pn) : nullptr;
    const double* ddv = dv ? io.in(dv, 2 * (size_t)p->N) : nullptr;
    const double* dvp = vp ? io.in(vp, z.ne) : nullptr;
    double* dvv = io.out(v, z.ne);
    if (io.status) return io.status;
    ILM_TRY(helmholtz_potentials_dev(p, dc, dd, ddv, p->g_a, p->g_b));
    ILM_TRY(launch_vecfield_from_potentials(p, p->g_a, p->g_b, dvp, dvp ? dvp + z.nu : nullptr, dvv, dvv + z.nu, deriv_div(p)));
    return io.finish();
}

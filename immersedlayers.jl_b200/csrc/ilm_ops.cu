// ilm_ops.cu -- surface <-> grid transfer kernels and staggered-grid stencils.
//
// regularize! / interpolate! and their normal-weighted variants
// (src/surface_operators.jl:13-343) and divergence!/grad!/curl!/laplacian!
// (src/grid_operators.jl:25-134 over CartesianGrids' stencils).
//   * regularization is a deterministic gather over the active cells (no
//     atomics): each cell adds its bucketed points in ascending point order
//     with separately rounded multiply and add -- the summation order of the
//     reference's CSC mat-vec -- so the field is bit-reproducible;
//   * interpolation uses one half-warp per surface point: 16 gathers and a
//     fixed shuffle tree;
//   * stencils are coalesced sweeps, x fastest, 8 rows per thread, with the
//     zero-fill, the difference and the 1/dx scaling fused into one pass
//     (the reference does fill!, stencil, ./= dx as three sweeps).
#include "ilm_internal.h"

namespace ilm {

#define ILM_LAUNCHED(p)               \
    do {                              \
        ILM_CUDA(cudaGetLastError()); \
        (p)->launches++;              \
    } while (0)

// a / d, correctly rounded, for the sweeps that divide every output by the grid spacing (_scale_derivative!,
// src/grid_operators.jl:8-9): with r = RN(1/d) computed once per thread, q0 = RN(a r), e = a - d q0 (exact in one
// FMA), q = RN(q0 + e r) is RN(a/d) (Markstein's theorem; it needs the significand of d not to be all ones and the
// quotient to stay in the normal range -- anything else takes the hardware division).  3 FP64 instructions instead of
// the ~18 of a division; the stencil outputs stay bit-identical to the oracle's `/ dx`.
struct ExactDiv {
    double d, r;
    bool fast;
    __device__ __forceinline__ explicit ExactDiv(double div) : d(div), r(1.0 / div) {
        const unsigned long long m = (unsigned long long)__double_as_longlong(div) & 0xFFFFFFFFFFFFFull;
        fast = m != 0xFFFFFFFFFFFFFull && isfinite(r) && r != 0.0;
    }
    __device__ __forceinline__ double operator()(double a) const {
        const double q0 = __dmul_rn(a, r);
        const double q1 = __fma_rn(__fma_rn(-d, q0, a), r, q0);
        const double m = fabs(q1);
        if (fast && ((m > 1e-290 && m < 1e290) || a == 0.0)) return a == 0.0 ? q0 : q1;
        return a / d;
    }
};

// ------------------------------------------------------------------ fill / scale
__global__ void k_fill(double* __restrict__ dst, size_t n, double value) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n2 = n >> 1;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        double2* d2 = reinterpret_cast<double2*>(dst);
        const double2 v2 = make_double2(value, value);
        for (size_t q = i; q < n2; q += stride) d2[q] = v2;
        if (i == 0 && (n & 1)) dst[n - 1] = value;
    } else {
        for (size_t q = i; q < n; q += stride) dst[q] = value;
    }
}

int launch_fill(ilm_plan* p, double* dst, size_t n, double value) {
    if (n == 0) return ILM_OK;
    size_t blocks = (n / 2 + 255) / 256;
    const size_t cap = (size_t)p->nsm * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_fill<<<(unsigned)blocks, 256, 0, p->stream>>>(dst, n, value);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_scale(double* __restrict__ w, size_t n, double s) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) w[i] = w[i] * s;
}
int launch_scale(ilm_plan* p, double* w, size_t n, double s) {
    if (n == 0) return ILM_OK;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)p->nsm * 16;
    if (blocks > cap) blocks = cap;
    k_scale<<<(unsigned)blocks, 256, 0, p->stream>>>(w, n, s);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_scale_store(const double* __restrict__ src, double* __restrict__ dst, int n, double scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = scale * src[i];
}
int launch_scale_store_column(ilm_plan* p, const double* src, double* dst, int n, double scale) {
    if (n == 0) return ILM_OK;
    k_scale_store<<<(n + 127) / 128, 128, 0, p->stream>>>(src, dst, n, scale);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// ------------------------------------------------------------------ regularize
__global__ void k_regularize(int ncell, int W2, const int* __restrict__ cell_idx, const int* __restrict__ cell_off,
                             const int* __restrict__ ent, const double* __restrict__ wR,
                             const double* __restrict__ f, const double* __restrict__ mul, double sign,
                             double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double sum = 0.0;
    const int q1 = cell_off[c + 1];
    for (int q = cell_off[c]; q < q1; ++q) {
        const int id = ent[q];
        const int k = id / W2;
        double v = f[k];
        if (mul) v = __dmul_rn(mul[k], v);
        v = sign < 0 ? -v : v;
        sum = __dadd_rn(sum, __dmul_rn(wR[id], v));
    }
    out[cell_idx[c]] = sum;
}

// forcing (src/forcing.jl:475-478, 487-490): the reference regularizes into a zero-filled scratch field and adds
// the whole field to dy; only the active cells can change, so the sum of each active cell is added in place
__global__ void k_regularize_add(int ncell, int W2, const int* __restrict__ cell_idx, const int* __restrict__ cell_off,
                                 const int* __restrict__ ent, const double* __restrict__ wR,
                                 const double* __restrict__ f, double* __restrict__ dy, int negate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double sum = 0.0;
    const int q1 = cell_off[c + 1];
    for (int q = cell_off[c]; q < q1; ++q) {
        const int id = ent[q];
        sum = __dadd_rn(sum, __dmul_rn(wR[id], f[id / W2]));
    }
    const int cell = cell_idx[c];
    dy[cell] = __dadd_rn(dy[cell], negate ? -sum : sum);
}

int launch_regularize_add(ilm_plan* p, const DevTable& t, const double* f, double* dy, bool negate) {
    if (t.ncell == 0) return ILM_OK;
    k_regularize_add<<<(t.ncell + 127) / 128, 128, 0, p->stream>>>(t.ncell, t.W * t.W, t.cell_idx, t.cell_off, t.ent, t.wR, f, dy,
                                                                    negate ? 1 : 0);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// area forcing (src/forcing.jl:463): dy .+= str .* mask (mask == nullptr: the whole domain, dy .+= str)
__global__ void k_forcing_area(const double* __restrict__ str, const double* __restrict__ mask, double* __restrict__ dy,
                               size_t n, int vec2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec2) {
        const size_t n2 = n >> 1;
        const double2* s2 = reinterpret_cast<const double2*>(str);
        const double2* m2 = reinterpret_cast<const double2*>(mask);
        double2* d2 = reinterpret_cast<double2*>(dy);
        for (size_t q = i; q < n2; q += stride) {
            double2 s = s2[q], d = d2[q];
            if (mask) { const double2 m = m2[q]; s.x = __dmul_rn(s.x, m.x); s.y = __dmul_rn(s.y, m.y); }
            d.x = __dadd_rn(d.x, s.x);
            d.y = __dadd_rn(d.y, s.y);
            d2[q] = d;
        }
        if (i == 0 && (n & 1)) dy[n - 1] = __dadd_rn(dy[n - 1], mask ? __dmul_rn(str[n - 1], mask[n - 1]) : str[n - 1]);
    } else {
        for (size_t q = i; q < n; q += stride) dy[q] = __dadd_rn(dy[q], mask ? __dmul_rn(str[q], mask[q]) : str[q]);
    }
}

int launch_forcing_area(ilm_plan* p, const double* str, const double* mask, double* dy, size_t n) {
    if (n == 0) return ILM_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(str) | reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(dy);
    const int vec2 = (al & 15) == 0;
    size_t blocks = ((vec2 ? n / 2 : n) + 255) / 256;
    const size_t cap = (size_t)p->nsm * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_forcing_area<<<(unsigned)blocks, 256, 0, p->stream>>>(str, mask, dy, n, vec2);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// regularize! = fill!(s, 0) + R f (src/surface_operators.jl:38-41) in ONE sweep: a block zero-fills its slice of the
// field, then (after a block barrier, which orders the two stores of a cell) gathers the active cells of that slice.
// The cell list is ascending, so the slice's cells are one contiguous run found by a binary search.
__global__ void __launch_bounds__(256) k_regularize_fused(size_t n, size_t chunk, int ncell, int W2, const int* __restrict__ cell_idx,
                                                          const int* __restrict__ cell_off, const int* __restrict__ ent,
                                                          const double* __restrict__ wR, const double* __restrict__ f,
                                                          const double* __restrict__ mul, double sign, double* __restrict__ out) {
    const size_t lo = (size_t)blockIdx.x * chunk;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    if (lo >= hi) return;
    if ((reinterpret_cast<uintptr_t>(out + lo) & 15) == 0) {
        double2* d2 = reinterpret_cast<double2*>(out + lo);
        const size_t n2 = (hi - lo) >> 1;
        for (size_t q = threadIdx.x; q < n2; q += 256) d2[q] = make_double2(0.0, 0.0);
        if (threadIdx.x == 0 && ((hi - lo) & 1)) out[hi - 1] = 0.0;
    } else {
        for (size_t q = lo + threadIdx.x; q < hi; q += 256) out[q] = 0.0;
    }
    int a = 0, b = ncell;                                   // first active cell >= lo
    while (a < b) { const int m = (a + b) >> 1; if ((size_t)cell_idx[m] < lo) a = m + 1; else b = m; }
    __syncthreads();
    for (int c = a + threadIdx.x; c < ncell; c += 256) {
        const size_t cell = (size_t)cell_idx[c];
        if (cell >= hi) break;
        double sum = 0.0;
        const int q1 = cell_off[c + 1];
        for (int q = cell_off[c]; q < q1; ++q) {
            const int id = ent[q];
            double v = f[id / W2];
            if (mul) v = __dmul_rn(mul[id / W2], v);
            v = sign < 0 ? -v : v;
            sum = __dadd_rn(sum, __dmul_rn(wR[id], v));
        }
        out[cell] = sum;
    }
}

int launch_regularize(ilm_plan* p, const DevTable& t, const double* f, const double* mul, double sign, double* out,
                      bool zero) {
    if (zero) {
        const size_t n = (size_t)t.mx * t.my;
        if (n == 0) return ILM_OK;
        size_t blocks = (size_t)p->nsm * 8;
        size_t chunk = ((n + blocks - 1) / blocks + 1) & ~(size_t)1;
        if (chunk < 512) chunk = 512;
        blocks = (n + chunk - 1) / chunk;
        k_regularize_fused<<<(unsigned)blocks, 256, 0, p->stream>>>(n, chunk, t.ncell, t.W * t.W, t.cell_idx, t.cell_off, t.ent,
                                                                    t.wR, f, mul, sign, out);
        ILM_LAUNCHED(p);
        return ILM_OK;
    }
    if (t.ncell == 0) return ILM_OK;
    k_regularize<<<(t.ncell + 127) / 128, 128, 0, p->stream>>>(t.ncell, t.W * t.W, t.cell_idx, t.cell_off, t.ent,
                                                               t.wR, f, mul, sign, out);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// unit-vector probe of the Schur builders: out = R e_col restricted to rows [rlo, rhi)
__global__ void k_regularize_unit(int W, int mx, int my, const int* __restrict__ i0, const int* __restrict__ j0,
                                  const double* __restrict__ wR, int col, double* __restrict__ out) {
    const int slot = threadIdx.x;
    if (slot >= W * W) return;
    const int a = slot % W, b = slot / W;
    const int i = i0[col] + a, j = j0[col] + b;
    if (i >= 0 && i < mx && j >= 0 && j < my) out[(size_t)j * mx + i] = wR[(size_t)col * W * W + slot];
}
int launch_regularize_unit(ilm_plan* p, const DevTable& t, int col, double* out, int rlo, int rhi) {
    ILM_TRY(launch_fill(p, out + (size_t)rlo * t.mx, (size_t)(rhi - rlo) * t.mx, 0.0));
    k_regularize_unit<<<1, 32, 0, p->stream>>>(t.W, t.mx, t.my, t.i0, t.j0, t.wR, col, out);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_rowsum(int ncell, const int* __restrict__ cell_off, const int* __restrict__ ent,
                         const double* __restrict__ wR, double* __restrict__ rowsum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double sum = 0.0;
    for (int q = cell_off[c]; q < cell_off[c + 1]; ++q) sum += wR[ent[q]];
    rowsum[c] = sum == 0.0 ? 1.0 : sum;
}
int launch_filter_rowsum(ilm_plan* p, DevTable& t) {
    if (t.ncell == 0) return ILM_OK;
    k_rowsum<<<(t.ncell + 127) / 128, 128, 0, p->stream>>>(t.ncell, t.cell_off, t.ent, t.wR, t.rowsum);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// C = Etilde R, one thread per row k (rows are independent -> no atomics)
__global__ void k_surface_filter(int N, int W, int mx, int my, const int* __restrict__ i0, const int* __restrict__ j0,
                                 const double* __restrict__ wE, const double* __restrict__ wR, int ncell,
                                 const int* __restrict__ cell_idx, const int* __restrict__ cell_off,
                                 const int* __restrict__ ent, const double* __restrict__ rowsum,
                                 double* __restrict__ C) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const int W2 = W * W;
    for (int b = 0; b < W; ++b)
        for (int a = 0; a < W; ++a) {
            const int i = i0[k] + a, j = j0[k] + b;
            if (i < 0 || i >= mx || j < 0 || j >= my) continue;
            const int lin = i + mx * j;
            int lo = 0, hi = ncell - 1, c = -1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int v = cell_idx[mid];
                if (v == lin) { c = mid; break; }
                if (v < lin) lo = mid + 1; else hi = mid - 1;
            }
            if (c < 0) continue;
            const double ef = wE[(size_t)k * W2 + b * W + a] / rowsum[c];
            for (int q = cell_off[c]; q < cell_off[c + 1]; ++q) {
                const int id = ent[q];
                const int l = id / W2;
                C[(size_t)l * N + k] += ef * wR[id];
            }
        }
}
int launch_surface_filter(ilm_plan* p, const DevTable& t, double* C) {
    const int N = p->N;
    ILM_TRY(launch_fill(p, C, (size_t)N * N, 0.0));
    if (N == 0) return ILM_OK;
    k_surface_filter<<<(N + 63) / 64, 64, 0, p->stream>>>(N, t.W, t.mx, t.my, t.i0, t.j0, t.wE, t.wR, t.ncell,
                                                           t.cell_idx, t.cell_off, t.ent, t.rowsum, C);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// ------------------------------------------------------------------ interpolate
struct TabView {
    int W, mx, my;
    const int *i0, *j0;
    const double* wE;
};
__device__ __forceinline__ double gather16(const TabView& t, const double* __restrict__ field, int k, int slot,
                                           bool live) {
    double val = 0.0;
    const int W2 = t.W * t.W;
    if (live && slot < W2) {
        const int a = slot % t.W, b = slot / t.W;
        const int i = t.i0[k] + a, j = t.j0[k] + b;
        if (i >= 0 && i < t.mx && j >= 0 && j < t.my)
            val = __dmul_rn(t.wE[(size_t)k * W2 + slot], field[(size_t)j * t.mx + i]);
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) val += __shfl_down_sync(0xffffffffu, val, off, 16);
    return val;
}

__global__ void k_interpolate(int N, TabView t, const double* __restrict__ field, double* __restrict__ f) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double v = gather16(t, field, k, slot, k < N);
    if (slot == 0 && k < N) f[k] = v;
}

static TabView view(const DevTable& t) { return TabView{t.W, t.mx, t.my, t.i0, t.j0, t.wE}; }

int launch_interpolate(ilm_plan* p, const DevTable& t, const double* field, double* f) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16;
    k_interpolate<<<(threads + 127) / 128, 128, 0, p->stream>>>(p->N, view(t), field, f);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_normal_interpolate(int N, TabView tu, TabView tv, const double* __restrict__ u,
                                     const double* __restrict__ v, const double* __restrict__ nx,
                                     const double* __restrict__ ny, int mode, double div, double* __restrict__ f) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double su = gather16(tu, u, k, slot, k < N);
    const double sv = gather16(tv, v, k, slot, k < N);
    if (slot == 0 && k < N) {
        double r;
        if (mode == ILM_NORMAL) r = __dadd_rn(__dmul_rn(nx[k], su), __dmul_rn(ny[k], sv));
        else r = __dsub_rn(__dmul_rn(nx[k], sv), __dmul_rn(ny[k], su));
        f[k] = r / div;
    }
}

// Fused  n . E_f (G phi)  /  n . E_f (C s): the edge value is formed on the fly from the node field
// (same difference, same multiply as stencil-then-gather: bit-identical, one full-field sweep less).
// OP 0: grad of Nodes{Primal}; OP 1: curl of Nodes{Dual}.  COMP 0: u entries, 1: v entries.
template <int OP, int COMP>
__device__ __forceinline__ double gather16_edge(const TabView& t, const double* __restrict__ nf, int NX, int NY, int k,
                                                int slot, bool live) {
    double val = 0.0;
    const int W2 = t.W * t.W;
    if (live && slot < W2) {
        const int a = slot % t.W, b = slot / t.W;
        const int i = t.i0[k] + a, j = t.j0[k] + b;
        if (i >= 0 && i < t.mx && j >= 0 && j < t.my) {
            double ev = 0.0;
            const int mp = NX - 1;
            if (OP == 0) {
                if (COMP == 0) { if (i >= 1 && i <= NX - 2) ev = nf[(size_t)j * mp + i] - nf[(size_t)j * mp + i - 1]; }
                else { if (j >= 1 && j <= NY - 2) ev = nf[(size_t)j * mp + i] - nf[(size_t)(j - 1) * mp + i]; }
            } else {
                if (COMP == 0) ev = nf[(size_t)(j + 1) * NX + i] - nf[(size_t)j * NX + i];
                else ev = nf[(size_t)j * NX + i] - nf[(size_t)j * NX + i + 1];
            }
            val = __dmul_rn(t.wE[(size_t)k * W2 + slot], ev);
        }
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) val += __shfl_down_sync(0xffffffffu, val, off, 16);
    return val;
}

template <int OP>
__global__ void k_normal_interpolate_fused(int N, TabView tu, TabView tv, const double* __restrict__ nf, int NX, int NY,
                                           const double* __restrict__ nx, const double* __restrict__ ny, int mode,
                                           double div, double* __restrict__ f) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double su = gather16_edge<OP, 0>(tu, nf, NX, NY, k, slot, k < N);
    const double sv = gather16_edge<OP, 1>(tv, nf, NX, NY, k, slot, k < N);
    if (slot == 0 && k < N) {
        double r;
        if (mode == ILM_NORMAL) r = __dadd_rn(__dmul_rn(nx[k], su), __dmul_rn(ny[k], sv));
        else r = __dsub_rn(__dmul_rn(nx[k], sv), __dmul_rn(ny[k], su));
        f[k] = r / div;
    }
}

// op 0: f = n . E_f G phi / div (phi: Nodes{Primal}); op 1: f = n . E_f C s / div (s: Nodes{Dual})
int launch_normal_interpolate_fused(ilm_plan* p, int op, int mode, const double* nodes, double* f, double div) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16, blocks = (threads + 127) / 128;
    if (op == 0)
        k_normal_interpolate_fused<0><<<blocks, 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]),
                                                                    nodes, p->g.NX, p->g.NY, p->nx, p->ny, mode, div, f);
    else
        k_normal_interpolate_fused<1><<<blocks, 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]),
                                                                    nodes, p->g.NX, p->g.NY, p->nx, p->ny, mode, div, f);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// unit-vector probe of regularize_normal!: rows [flo, fhi) of both Edges components are zeroed and the
// two WxW patches n_u R_u e_col, n_v R_v e_col written (mode NORMAL: (n.u, n.v); CROSS: (n.v, -n.u))
__global__ void k_regularize_normal_unit(int W, int NX, int NY, const int* __restrict__ iu, const int* __restrict__ ju,
                                         const double* __restrict__ wu, const int* __restrict__ iv,
                                         const int* __restrict__ jv, const double* __restrict__ wv,
                                         const double* __restrict__ nx, const double* __restrict__ ny, int mode, int col,
                                         double* __restrict__ u, double* __restrict__ v) {
    const int slot = threadIdx.x & 15, comp = threadIdx.x >> 4;
    if (slot >= W * W || comp > 1) return;
    const int a = slot % W, b = slot / W;
    const double fu = mode == ILM_NORMAL ? nx[col] : ny[col];
    const double fv = mode == ILM_NORMAL ? ny[col] : -nx[col];
    if (comp == 0) {
        const int i = iu[col] + a, j = ju[col] + b;
        if (i >= 0 && i < NX && j >= 0 && j < NY - 1) u[(size_t)j * NX + i] = __dmul_rn(wu[(size_t)col * W * W + slot], fu);
    } else {
        const int i = iv[col] + a, j = jv[col] + b;
        if (i >= 0 && i < NX - 1 && j >= 0 && j < NY) v[(size_t)j * (NX - 1) + i] = __dmul_rn(wv[(size_t)col * W * W + slot], fv);
    }
}
int launch_regularize_normal_unit(ilm_plan* p, int mode, int col, double* u, double* v, int flo, int fhi, bool fill) {
    const DevTable& tu = p->tab[ILM_XEDGES];
    const DevTable& tv = p->tab[ILM_YEDGES];
    if (fill) {
        const int ulo = max(flo, 0), uhi = min(fhi, tu.my), vlo = max(flo, 0), vhi = min(fhi, tv.my);
        if (uhi > ulo) ILM_TRY(launch_fill(p, u + (size_t)ulo * tu.mx, (size_t)(uhi - ulo) * tu.mx, 0.0));
        if (vhi > vlo) ILM_TRY(launch_fill(p, v + (size_t)vlo * tv.mx, (size_t)(vhi - vlo) * tv.mx, 0.0));
    }
    k_regularize_normal_unit<<<1, 32, 0, p->stream>>>(tu.W, p->g.NX, p->g.NY, tu.i0, tu.j0, tu.wR, tv.i0, tv.j0, tv.wR, p->nx,
                                                      p->ny, mode, col, u, v);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

int launch_normal_interpolate(ilm_plan* p, int mode, const double* u, const double* v, double* f, double div) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16;
    k_normal_interpolate<<<(threads + 127) / 128, 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]),
                                                                       view(p->tab[ILM_YEDGES]), u, v, p->nx, p->ny,
                                                                       mode, div, f);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// ------------------------------------------------------------------ stencils
// thread = one x, ROWS consecutive y; grid covers the NX x NY superset.
// block = ST_BX consecutive x, each thread walks ST_ROWS rows.  Swept under ncu on B200 at 4096^2 (128..1024 x 4..16,
// profiles/r2_stencil_geometry.txt): 256 x 4 is the best or within noise of it for every sweep; the one-read / two-write
// sweeps (grad, curl nodes -> edges) stay at 70-75 % of the copy bandwidth for every geometry.
#ifndef ILM_ST_BX
#define ILM_ST_BX 256
#endif
#ifndef ILM_ST_ROWS
#define ILM_ST_ROWS 4
#endif
constexpr int ST_BX = ILM_ST_BX, ST_ROWS = ILM_ST_ROWS;
static dim3 st_grid(int NX, int NY) { return dim3((NX + ST_BX - 1) / ST_BX, (NY + ST_ROWS - 1) / ST_ROWS); }

// p[x,y] = -u[x,y] + u[x+1,y] - v[x,y] + v[x,y+1]          (A.3)
__global__ void k_divergence(int NX, int NY, const double* __restrict__ u, const double* __restrict__ v,
                             double* __restrict__ out, double div, int ybeg, int yend) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = ybeg + blockIdx.y * ST_ROWS;
    if (x >= NX - 1) return;
    const int mxv = NX - 1;
    double vprev = (y0 < NY) ? v[(size_t)y0 * mxv + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= NY - 1 || y >= yend) break;
        const double vn = v[(size_t)(y + 1) * mxv + x];
        const double ul = u[(size_t)y * NX + x], ur = u[(size_t)y * NX + x + 1];
        out[(size_t)y * mxv + x] = dv(((-ul + ur) - vprev) + vn);
        vprev = vn;
    }
}
int launch_divergence(ilm_plan* p, const double* u, const double* v, double* out, double div, int ybeg, int yend) {
    if (ybeg < 0) { ybeg = 0; yend = p->g.NY; }
    k_divergence<<<st_grid(p->g.NX, yend - ybeg), ST_BX, 0, p->stream>>>(p->g.NX, p->g.NY, u, v, out, div, ybeg, yend);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// u[x,y] = p[x,y]-p[x-1,y] (x in 2:NX-1), v[x,y] = p[x,y]-p[x,y-1] (y in 2:NY-1), zero elsewhere
__global__ void k_grad(int NX, int NY, const double* __restrict__ pn, double* __restrict__ u, double* __restrict__ v,
                       double div) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = blockIdx.y * ST_ROWS;
    if (x >= NX) return;
    const int mp = NX - 1;
    double pprev = (y0 >= 1 && y0 - 1 < NY - 1 && x < mp) ? pn[(size_t)(y0 - 1) * mp + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= NY) break;
        const bool prow = y < NY - 1;
        const double pc = (prow && x < mp) ? pn[(size_t)y * mp + x] : 0.0;
        if (prow) {   // u row
            double val = 0.0;
            if (x >= 1 && x <= NX - 2) val = dv(pc - pn[(size_t)y * mp + x - 1]);
            u[(size_t)y * NX + x] = val;
        }
        if (x < mp) {  // v row
            double val = 0.0;
            if (y >= 1 && y <= NY - 2) val = dv(pc - pprev);
            v[(size_t)y * mp + x] = val;
        }
        pprev = pc;
    }
}
int launch_grad(ilm_plan* p, const double* in, double* u, double* v, double div) {
    k_grad<<<st_grid(p->g.NX, p->g.NY), ST_BX, 0, p->stream>>>(p->g.NX, p->g.NY, in, u, v, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// u[x,y] = s[x,y+1]-s[x,y] ; v[x,y] = s[x,y]-s[x+1,y]
__global__ void k_curl_n2e(int NX, int NY, const double* __restrict__ s, double* __restrict__ u,
                           double* __restrict__ v, double div) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = blockIdx.y * ST_ROWS;
    if (x >= NX) return;
    const int mxv = NX - 1;
    double sc = (y0 < NY) ? s[(size_t)y0 * NX + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= NY) break;
        if (x < mxv) v[(size_t)y * mxv + x] = dv(sc - s[(size_t)y * NX + x + 1]);
        if (y < NY - 1) {
            const double sn = s[(size_t)(y + 1) * NX + x];
            u[(size_t)y * NX + x] = dv(sn - sc);
            sc = sn;
        }
    }
}
int launch_curl_n2e(ilm_plan* p, const double* s, double* u, double* v, double div) {
    k_curl_n2e<<<st_grid(p->g.NX, p->g.NY), ST_BX, 0, p->stream>>>(p->g.NX, p->g.NY, s, u, v, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// Helmholtz recomposition (src/helmholtz.jl:285-307): v = curl(psi) + grad(phi) [+ vp] in one sweep; every term is
// rounded as in k_curl_n2e / k_grad and the sums are taken in the reference's order (v .= vpsi .+ vphi; v .+= vp)
__global__ void k_vecfield_from_potentials(int NX, int NY, const double* __restrict__ psi, const double* __restrict__ phi,
                                           const double* __restrict__ vpu, const double* __restrict__ vpv,
                                           double* __restrict__ u, double* __restrict__ v, double div) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = blockIdx.y * ST_ROWS;
    if (x >= NX) return;
    const int mp = NX - 1;
    double sc = (psi && y0 < NY) ? psi[(size_t)y0 * NX + x] : 0.0;
    double pprev = (phi && y0 >= 1 && y0 - 1 < NY - 1 && x < mp) ? phi[(size_t)(y0 - 1) * mp + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= NY) break;
        const bool prow = y < NY - 1;
        const double pc = (phi && prow && x < mp) ? phi[(size_t)y * mp + x] : 0.0;
        if (x < mp) {   // v row
            const double cv = psi ? dv(sc - psi[(size_t)y * NX + x + 1]) : 0.0;
            double gv = 0.0;
            if (phi && y >= 1 && y <= NY - 2) gv = dv(pc - pprev);
            double val = __dadd_rn(cv, gv);
            if (vpv) val = __dadd_rn(val, vpv[(size_t)y * mp + x]);
            v[(size_t)y * mp + x] = val;
        }
        if (prow) {     // u row
            double cu = 0.0;
            if (psi) {
                const double sn = psi[(size_t)(y + 1) * NX + x];
                cu = dv(sn - sc);
                sc = sn;
            }
            double gu = 0.0;
            if (phi && x >= 1 && x <= NX - 2) gu = dv(pc - phi[(size_t)y * mp + x - 1]);
            double val = __dadd_rn(cu, gu);
            if (vpu) val = __dadd_rn(val, vpu[(size_t)y * NX + x]);
            u[(size_t)y * NX + x] = val;
        }
        pprev = pc;
    }
}
int launch_vecfield_from_potentials(ilm_plan* p, const double* psi, const double* phi, const double* vpu, const double* vpv,
                                    double* u, double* v, double div) {
    k_vecfield_from_potentials<<<st_grid(p->g.NX, p->g.NY), ST_BX, 0, p->stream>>>(p->g.NX, p->g.NY, psi, phi, vpu, vpv, u, v, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// w[x,y] = u[x,y-1]-u[x,y]-v[x-1,y]+v[x,y], x in 2:NX-1, y in 2:NY-1, zero elsewhere
__global__ void k_curl_e2n(int NX, int NY, const double* __restrict__ u, const double* __restrict__ v,
                           double* __restrict__ w, double div, int ybeg, int yend) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = ybeg + blockIdx.y * ST_ROWS;
    if (x >= NX) return;
    const int mxv = NX - 1;
    double uprev = (y0 >= 1 && y0 - 1 < NY - 1) ? u[(size_t)(y0 - 1) * NX + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= NY || y >= yend) break;
        const double uc = (y < NY - 1) ? u[(size_t)y * NX + x] : 0.0;
        double val = 0.0;
        if (x >= 1 && x <= NX - 2 && y >= 1 && y <= NY - 2)
            val = dv(((uprev - uc) - v[(size_t)y * mxv + x - 1]) + v[(size_t)y * mxv + x]);
        w[(size_t)y * NX + x] = val;
        uprev = uc;
    }
}
int launch_curl_e2n(ilm_plan* p, const double* u, const double* v, double* w, double div, int ybeg, int yend) {
    if (ybeg < 0) { ybeg = 0; yend = p->g.NY; }
    k_curl_e2n<<<st_grid(p->g.NX, yend - ybeg), ST_BX, 0, p->stream>>>(p->g.NX, p->g.NY, u, v, w, div, ybeg, yend);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// 5-point Laplacian times factor on the interior of an mx x my field, zero on its border
__global__ void k_laplacian(int mx, int my, const double* __restrict__ in, double* __restrict__ out, double factor) {
    const int x = blockIdx.x * ST_BX + threadIdx.x;
    const int y0 = blockIdx.y * ST_ROWS;
    if (x >= mx) return;
    double below = (y0 >= 1) ? in[(size_t)(y0 - 1) * mx + x] : 0.0;
    double cur = (y0 < my) ? in[(size_t)y0 * mx + x] : 0.0;
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) {
        const int y = y0 + r;
        if (y >= my) break;
        const double above = (y + 1 < my) ? in[(size_t)(y + 1) * mx + x] : 0.0;
        double val = 0.0;
        if (x >= 1 && x <= mx - 2 && y >= 1 && y <= my - 2) {
            const double l = in[(size_t)y * mx + x - 1], rr = in[(size_t)y * mx + x + 1];
            val = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(-4.0, cur), l), rr), below), above), factor);
        }
        out[(size_t)y * mx + x] = val;
        below = cur;
        cur = above;
    }
}
int launch_laplacian(ilm_plan* p, const double* in, double* out, int mx, int my, double factor) {
    k_laplacian<<<st_grid(mx, my), ST_BX, 0, p->stream>>>(mx, my, in, out, factor);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// ------------------------------------------------------------------ LGF multiplier source
// h[i,j] = eps_i eps_j (G[i,j] - c0), eps_0 = 1, eps_{>0} = 2   (see ilm_conv.cuh)
__global__ void k_lgf_prep(const double* __restrict__ table, int ld, int NX, int NY, double c0,
                           double* __restrict__ h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= NX || j >= NY) return;
    const double e = (i ? 2.0 : 1.0) * (j ? 2.0 : 1.0);
    h[(size_t)j * NX + i] = e * (table[(size_t)j * ld + i] - c0);
}
int launch_lgf_prep(ilm_plan* p, const double* table, int ld, int NX, int NY, double c0, double* h) {
    k_lgf_prep<<<dim3((NX + 127) / 128, NY), 128, 0, p->stream>>>(table, ld, NX, NY, c0, h);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

}  // namespace ilm

// =====================================================================================
// vector-cache operators: TensorData / EdgeGradient forms (src/surface_operators.jl:113-138,
// 172-215, 242-343) and the Edges <-> EdgeGradient stencils (src/grid_operators.jl:61-71).
// Tensor convention (parity unpinned): component (i,j) = (velocity
// component, derivative direction): [dudx, dudy, dvdx, dvdy]; n (x) v -> v_i n_j; n . T -> sum_j n_j T_ij.
// =====================================================================================
namespace ilm {

// T = n (x) v (mode 0) or n (x) v + (n (x) v)^T - (n.v) I (mode 1)
__global__ void k_tensor_from_vector(int N, const double* __restrict__ nx, const double* __restrict__ ny,
                                     const double* __restrict__ v, int mode, double* __restrict__ T) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double n1 = nx[k], n2 = ny[k], vu = v[k], vv = v[N + k];
    const double t0 = __dmul_rn(n1, vu), t1 = __dmul_rn(n2, vu), t2 = __dmul_rn(n1, vv), t3 = __dmul_rn(n2, vv);
    if (mode == 0) {
        T[k] = t0; T[N + k] = t1; T[2 * N + k] = t2; T[3 * N + k] = t3;
    } else {
        const double dot = __dadd_rn(t0, t3);
        T[k] = __dsub_rn(__dadd_rn(t0, t0), dot);
        T[N + k] = __dadd_rn(t1, t2);
        T[2 * N + k] = __dadd_rn(t2, t1);
        T[3 * N + k] = __dsub_rn(__dadd_rn(t3, t3), dot);
    }
}
int launch_tensor_from_vector(ilm_plan* p, int mode, const double* v, double* T) {
    if (p->N == 0) return ILM_OK;
    k_tensor_from_vector<<<(p->N + 127) / 128, 128, 0, p->stream>>>(p->N, p->nx, p->ny, v, mode, T);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// t = n . S (mode 0) or n . (S + S^T - tr(S) I) (mode 1), then / div
__global__ void k_tensor_dot(int N, const double* __restrict__ nx, const double* __restrict__ ny,
                             const double* __restrict__ S, int mode, double div, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double s0 = S[k], s1 = S[N + k], s2 = S[2 * N + k], s3 = S[3 * N + k];
    if (mode == 1) {
        const double tr = __dadd_rn(s0, s3), o = __dadd_rn(s2, s1);
        s0 = __dsub_rn(__dadd_rn(s0, s0), tr);
        s3 = __dsub_rn(__dadd_rn(s3, s3), tr);
        s1 = o; s2 = o;
    }
    const double n1 = nx[k], n2 = ny[k];
    out[k] = __dadd_rn(__dmul_rn(n1, s0), __dmul_rn(n2, s1)) / div;
    out[N + k] = __dadd_rn(__dmul_rn(n1, s2), __dmul_rn(n2, s3)) / div;
}
int launch_tensor_dot(ilm_plan* p, int mode, const double* S, double div, double* out) {
    if (p->N == 0) return ILM_OK;
    k_tensor_dot<<<(p->N + 127) / 128, 128, 0, p->stream>>>(p->N, p->nx, p->ny, S, mode, div, out);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// op 0: s = n.u v.v - n.v v.u ; 1: s = n.u v.u + n.v v.v ; 2: v = (n.v s, -n.u s) ; 3: v = (n.u s, n.v s)
__global__ void k_vec_pointwise(int N, const double* __restrict__ nx, const double* __restrict__ ny,
                                const double* __restrict__ in, int op, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double n1 = nx[k], n2 = ny[k];
    if (op == 0) out[k] = __dsub_rn(__dmul_rn(n1, in[N + k]), __dmul_rn(n2, in[k]));
    else if (op == 1) out[k] = __dadd_rn(__dmul_rn(n1, in[k]), __dmul_rn(n2, in[N + k]));
    else if (op == 2) { out[k] = __dmul_rn(n2, in[k]); out[N + k] = -__dmul_rn(n1, in[k]); }
    else { out[k] = __dmul_rn(n1, in[k]); out[N + k] = __dmul_rn(n2, in[k]); }
}
int launch_vec_pointwise(ilm_plan* p, int op, const double* in, double* out) {
    if (p->N == 0) return ILM_OK;
    k_vec_pointwise<<<(p->N + 127) / 128, 128, 0, p->stream>>>(p->N, p->nx, p->ny, in, op, out);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// grad!(EdgeGradient <- Edges): A.3
__global__ void k_grad_tensor(int NX, int NY, const double* __restrict__ u, const double* __restrict__ v,
                              double* __restrict__ dudx, double* __restrict__ dudy, double* __restrict__ dvdx,
                              double* __restrict__ dvdy, double div) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= NX || y >= NY) return;
    const int mp = NX - 1;
    if (x < NX - 1 && y < NY - 1) {
        dudx[(size_t)y * mp + x] = dv(u[(size_t)y * NX + x + 1] - u[(size_t)y * NX + x]);
        dvdy[(size_t)y * mp + x] = dv(v[(size_t)(y + 1) * mp + x] - v[(size_t)y * mp + x]);
    }
    double a = 0.0, b = 0.0;
    if (x >= 1 && x <= NX - 2 && y >= 1 && y <= NY - 2) {
        a = dv(u[(size_t)y * NX + x] - u[(size_t)(y - 1) * NX + x]);
        b = dv(v[(size_t)y * mp + x] - v[(size_t)y * mp + x - 1]);
    }
    dudy[(size_t)y * NX + x] = a;
    dvdx[(size_t)y * NX + x] = b;
}
int launch_grad_tensor(ilm_plan* p, const double* edges, double* eg, double div) {
    const int NX = p->g.NX, NY = p->g.NY;
    const size_t Pn = (size_t)(NX - 1) * (NY - 1), Pd = (size_t)NX * NY, nu = (size_t)NX * (NY - 1);
    k_grad_tensor<<<dim3((NX + 127) / 128, NY), 128, 0, p->stream>>>(NX, NY, edges, edges + nu, eg, eg + Pn, eg + Pn + Pd,
                                                                     eg + Pn + 2 * Pd, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// divergence!(Edges <- EdgeGradient): A.3
__global__ void k_div_tensor(int NX, int NY, const double* __restrict__ dudx, const double* __restrict__ dudy,
                             const double* __restrict__ dvdx, const double* __restrict__ dvdy,
                             double* __restrict__ u, double* __restrict__ v, double div) {
    const ExactDiv dv(div);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= NX || y >= NY) return;
    const int mp = NX - 1;
    if (y < NY - 1) {
        double val = 0.0;
        if (x >= 1 && x <= NX - 2)
            val = dv(((dudx[(size_t)y * mp + x] - dudx[(size_t)y * mp + x - 1]) + dudy[(size_t)(y + 1) * NX + x]) -
                   dudy[(size_t)y * NX + x]);
        u[(size_t)y * NX + x] = val;
    }
    if (x < NX - 1) {
        double val = 0.0;
        if (y >= 1 && y <= NY - 2)
            val = dv(((dvdx[(size_t)y * NX + x + 1] - dvdx[(size_t)y * NX + x]) + dvdy[(size_t)y * mp + x]) -
                   dvdy[(size_t)(y - 1) * mp + x]);
        v[(size_t)y * mp + x] = val;
    }
}
int launch_div_tensor(ilm_plan* p, const double* eg, double* edges, double div) {
    const int NX = p->g.NX, NY = p->g.NY;
    const size_t Pn = (size_t)(NX - 1) * (NY - 1), Pd = (size_t)NX * NY, nu = (size_t)NX * (NY - 1);
    k_div_tensor<<<dim3((NX + 127) / 128, NY), 128, 0, p->stream>>>(NX, NY, eg, eg + Pn, eg + Pn + Pd, eg + Pn + 2 * Pd,
                                                                    edges, edges + nu, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

}  // namespace ilm

// =====================================================================================
// Direct-table form of S = -scale * E L^-1 R: the response of L^-1 to a
// unit impulse is the (shifted) LGF table itself, so
//   S[k,l] = -(scale/factor) * sum_p E[k,p] * sum_q (G(|p-q|) - c0) * R[q,l]
// needs no transform at all (O(N^2 W^4) table look-ups).  Cross-check of the column-solve path
// and an optional fast builder; bench.py's metric path always uses the column solves.
// =====================================================================================
namespace ilm {

__global__ void __launch_bounds__(256)
k_schur_direct(int N, int W, int mx, const int* __restrict__ i0, const int* __restrict__ j0,
               const double* __restrict__ wE, const double* __restrict__ wR, const double* __restrict__ G, int ldg,
               int ng, int my, double c0, double coef, int col_begin, double* __restrict__ A) {
    const int l = col_begin + blockIdx.y;
    __shared__ double rw[16];
    __shared__ int ri[16], rj[16];
    const int W2 = W * W;
    if (threadIdx.x < W2) {
        const int a = threadIdx.x % W, b = threadIdx.x / W;
        const int i = i0[l] + a, j = j0[l] + b;
        const bool ok = i >= 0 && i < mx && j >= 0 && j < my;
        rw[threadIdx.x] = ok ? wR[(size_t)l * W2 + threadIdx.x] : 0.0;
        ri[threadIdx.x] = i; rj[threadIdx.x] = j;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double sum = 0.0;
    for (int b = 0; b < W; ++b)
        for (int a = 0; a < W; ++a) {
            const int pi = i0[k] + a, pj = j0[k] + b;
            if (pi < 0 || pi >= mx || pj < 0 || pj >= my) continue;
            double inner = 0.0;
            for (int q = 0; q < W2; ++q) {
                const int di = abs(pi - ri[q]), dj = abs(pj - rj[q]);
                const double gv = (di < ng && dj < ng) ? G[(size_t)dj * ldg + di] : 0.0;     // compact tables: 0 beyond
                inner += (gv - c0) * rw[q];
            }
            sum += wE[(size_t)k * W2 + b * W + a] * inner;
        }
    A[(size_t)blockIdx.y * N + k] = coef * sum;
}

int launch_schur_direct(ilm_plan* p, const double* G, int ldg, int ng, double c0, double factor, double scale, int col_begin,
                        int col_end, double* A) {
    const DevTable& t = p->tab[ILM_NODES_PRIMAL];
    const int N = p->N, ncols = col_end - col_begin;
    if (N == 0 || ncols == 0) return ILM_OK;
    dim3 grid((N + 255) / 256, ncols);
    k_schur_direct<<<grid, 256, 0, p->stream>>>(N, t.W, t.mx, t.i0, t.j0, t.wE, t.wR, G, ldg, ng, t.my, c0, -scale / factor,
                                                col_begin, A);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

}  // namespace ilm

// =====================================================================================
// Fused pre/post kernels of the create_RTLinvR probe (two columns per launch):
//   pre : rows [rlo, rhi) of the two right-hand sides = R e_c0, R e_c1 (zero elsewhere in the rows)
//   post: A[:, c] = coef * E field_c   for both fields
// (replace 2 fills + 2 scatters and 2 interpolations + 2 scale-stores per column pair)
// =====================================================================================
namespace ilm {

__global__ void k_probe_pre(int W, int mx, int my, const int* __restrict__ i0, const int* __restrict__ j0,
                            const double* __restrict__ wR, int col0, int ncol, double* __restrict__ g0,
                            double* __restrict__ g1, int rlo, int rhi) {
    const int q = blockIdx.y;                   // which column / field
    if (q >= ncol) return;
    const int col = col0 + q;
    double* out = q ? g1 : g0;
    const int ci = i0[col], cj = j0[col];
    const size_t n = (size_t)(rhi - rlo) * mx;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % mx), j = rlo + (int)(t / mx);
        const int a = i - ci, b = j - cj;
        double v = 0.0;
        if (a >= 0 && a < W && b >= 0 && b < W) v = wR[(size_t)col * W * W + b * W + a];
        out[(size_t)j * mx + i] = v;
    }
}
int launch_probe_pre(ilm_plan* p, const DevTable& t, int col0, int ncol, double* g0, double* g1, int rlo, int rhi) {
    const size_t n = (size_t)(rhi - rlo) * t.mx;
    int bx = (int)((n + 255) / 256);
    if (bx > 2 * p->nsm) bx = 2 * p->nsm;
    k_probe_pre<<<dim3(bx, ncol), 256, 0, p->stream>>>(t.W, t.mx, t.my, t.i0, t.j0, t.wR, col0, ncol, g0, g1, rlo, rhi);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_probe_post(int N, TabView t, const double* __restrict__ g0, const double* __restrict__ g1, double coef,
                             double* __restrict__ d0, double* __restrict__ d1) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double* field = blockIdx.y ? g1 : g0;
    double* dst = blockIdx.y ? d1 : d0;
    const double v = gather16(t, field, k, slot, k < N);
    if (slot == 0 && k < N) dst[k] = coef * v;
}
// column c (and c + 1) of the Schur complement from the window-row sums that pass C left in t.part
__global__ void k_probe_post_sum(int N, int W, int my, const int* __restrict__ j0, const double2* __restrict__ part, double coef,
                                 double* __restrict__ d0, double* __restrict__ d1) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double re = 0.0, im = 0.0;
    for (int b = 0; b < W; ++b) {
        const int j = j0[k] + b;
        if (j >= 0 && j < my) { const double2 v = part[k * W + b]; re += v.x; im += v.y; }
    }
    d0[k] = coef * re;
    if (d1) d1[k] = coef * im;
}
// the same for a batch of column pairs: pair y (blockIdx.y) left its row sums in slice y of t.part and owns the
// columns 2y, 2y + 1 of dA0 (ncols columns in all; the last pair may be a single column)
__global__ void k_probe_post_sum_batch(int N, int W, int my, const int* __restrict__ j0, const double2* __restrict__ part,
                                       size_t stride, double coef, double* __restrict__ dA0, int ncols, const int* __restrict__ pairolo) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const int pair = blockIdx.y;
    if (pairolo && max(j0[k], 0) < pairolo[pair]) return;        // rows below the pair's first inverted row: filled by symmetry
    const double2* pp = part + (size_t)pair * stride;
    double re = 0.0, im = 0.0;
    for (int b = 0; b < W; ++b) {
        const int j = j0[k] + b;
        if (j >= 0 && j < my) { const double2 v = pp[k * W + b]; re += v.x; im += v.y; }
    }
    dA0[(size_t)(2 * pair) * N + k] = coef * re;
    if (2 * pair + 1 < ncols) dA0[(size_t)(2 * pair + 1) * N + k] = coef * im;
}
int launch_probe_post_sum_batch(ilm_plan* p, const DevTable& t, int npairs, int ncols, double coef, double* dA0, const int* pairolo,
                                int slice0, cudaStream_t st) {
    if (npairs <= 0) return ILM_OK;
    if (!st) st = p->stream;
    k_probe_post_sum_batch<<<dim3((p->N + 127) / 128, npairs), 128, 0, st>>>(p->N, t.W, t.my, t.j0, t.part + (size_t)slice0 * t.part_stride,
                                                                             t.part_stride, coef, dA0, ncols, pairolo);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// Symmetric Schur build (create_RTLinvR: S[k,c] = coef wgt_c sum e_k K e_c with the same windows e on both sides): column c was
// probed over the rows from its own window upwards only, so S[k, c] is there for the points k whose windows start at or above
// that row; the others come from the mirror entry, S[k, c] = S[c, k] wgt_c / wgt_k.  Ssort holds the columns in sorted order.
__global__ void k_schur_symm_finish(int N, const double* __restrict__ Ssort, const int* __restrict__ pos, const int* __restrict__ cololo,
                                    const int* __restrict__ j0, const double* __restrict__ ds, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (k >= N) return;
    const int olo = cololo[c];
    const bool valid = olo < 0 || max(j0[k], 0) >= olo;
    double v;
    if (valid) v = Ssort[(size_t)pos[c] * N + k];
    else {
        v = Ssort[(size_t)pos[k] * N + c];
        if (ds) v = v * ds[c] / ds[k];
    }
    out[(size_t)c * N + k] = v;
}
int launch_schur_symm_finish(ilm_plan* p, const double* Ssort, const int* pos, const int* cololo, const int* j0, const double* ds, double* out) {
    if (p->N == 0) return ILM_OK;
    k_schur_symm_finish<<<dim3((p->N + 127) / 128, p->N), 128, 0, p->stream>>>(p->N, Ssort, pos, cololo, j0, ds, out);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

int launch_probe_post_sum(ilm_plan* p, const DevTable& t, int ncol, double coef, double* d0, double* d1) {
    k_probe_post_sum<<<(p->N + 127) / 128, 128, 0, p->stream>>>(p->N, t.W, t.my, t.j0, t.part, coef, d0, ncol > 1 ? d1 : nullptr);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

int launch_probe_post(ilm_plan* p, const DevTable& t, int ncol, const double* g0, const double* g1, double coef, double* d0,
                      double* d1) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16;
    k_probe_post<<<dim3((threads + 127) / 128, ncol), 128, 0, p->stream>>>(p->N, view(t), g0, g1, coef, d0, d1);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// Fused pre/post kernels of the ScalarData stencil probes (create_CLinvCT, create_GLinvD, create_GLinvD_cross on a
// scalar cache, src/matrix_operators.jl:40-61,135-155,195-215), two columns per launch:
//   pre : rows [rlo, rhi) of the right-hand side = curl / divergence of the two W x W edge patches
//         (n_u R_u e_c, n_v R_v e_c) (normal) or (n_v R_u e_c, -n_u R_v e_c) (cross), divided by dx
//   post: A[:, c] = coef * (n . E_f (stencil field)) / dx  (the fused gather of k_normal_interpolate_fused)
// Same operations in the same order as regularize_normal! -> stencil and stencil -> normal_interpolate! -> scale.
template <int OP>       // 0: divergence -> Nodes{Primal} (GLinvD), 1: curl -> Nodes{Dual} (CLinvCT)
__global__ void k_sprobe_pre(int NX, int NY, TabView tu, TabView tv, const double* __restrict__ wRu, const double* __restrict__ wRv,
                             const double* __restrict__ nx, const double* __restrict__ ny, int mode, int col0, int ncol,
                             double* __restrict__ g0, double* __restrict__ g1, int rlo, int rhi, double div) {
    const int q = blockIdx.y;
    if (q >= ncol) return;
    const int c = col0 + q, W = tu.W;
    double* out = q ? g1 : g0;
    const double fu = mode == ILM_NORMAL ? nx[c] : ny[c];
    const double fv = mode == ILM_NORMAL ? ny[c] : -nx[c];
    const int iu = tu.i0[c], ju = tu.j0[c], iv = tv.i0[c], jv = tv.j0[c];
    auto U = [&](int i, int j) -> double {
        const int a = i - iu, b = j - ju;
        if (a < 0 || a >= W || b < 0 || b >= W || i < 0 || i >= NX || j < 0 || j >= NY - 1) return 0.0;
        return __dmul_rn(wRu[(size_t)c * W * W + b * W + a], fu);
    };
    auto V = [&](int i, int j) -> double {
        const int a = i - iv, b = j - jv;
        if (a < 0 || a >= W || b < 0 || b >= W || i < 0 || i >= NX - 1 || j < 0 || j >= NY) return 0.0;
        return __dmul_rn(wRv[(size_t)c * W * W + b * W + a], fv);
    };
    const int mx = OP == 0 ? NX - 1 : NX, my = OP == 0 ? NY - 1 : NY;
    const int r1 = rhi < my ? rhi : my;
    if (r1 <= rlo) return;
    const size_t n = (size_t)(r1 - rlo) * mx;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % mx), j = rlo + (int)(idx / mx);
        double val;
        if (OP == 0) val = (((-U(i, j) + U(i + 1, j)) - V(i, j)) + V(i, j + 1)) / div;
        else val = (i >= 1 && i <= NX - 2 && j >= 1 && j <= NY - 2) ? (((U(i, j - 1) - U(i, j)) - V(i - 1, j)) + V(i, j)) / div : 0.0;
        out[(size_t)j * mx + i] = val;
    }
}
int launch_sprobe_pre(ilm_plan* p, int op, int mode, int col0, int ncol, double* g0, double* g1, int rlo, int rhi, double div) {
    const DevTable& tu = p->tab[ILM_XEDGES];
    const DevTable& tv = p->tab[ILM_YEDGES];
    const size_t n = (size_t)(rhi - rlo) * p->g.NX;
    int bx = (int)((n + 255) / 256);
    if (bx > 2 * p->nsm) bx = 2 * p->nsm;
    if (bx < 1) bx = 1;
    if (op == 0)
        k_sprobe_pre<0><<<dim3(bx, ncol), 256, 0, p->stream>>>(p->g.NX, p->g.NY, view(tu), view(tv), tu.wR, tv.wR, p->nx, p->ny, mode,
                                                             col0, ncol, g0, g1, rlo, rhi, div);
    else
        k_sprobe_pre<1><<<dim3(bx, ncol), 256, 0, p->stream>>>(p->g.NX, p->g.NY, view(tu), view(tv), tu.wR, tv.wR, p->nx, p->ny, mode,
                                                             col0, ncol, g0, g1, rlo, rhi, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

template <int OP>
__global__ void k_sprobe_post(int N, TabView tu, TabView tv, int NX, int NY, const double* __restrict__ g0,
                              const double* __restrict__ g1, const double* __restrict__ nx, const double* __restrict__ ny, int mode,
                              double div, double coef, double* __restrict__ d0, double* __restrict__ d1) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double* nf = blockIdx.y ? g1 : g0;
    double* dst = blockIdx.y ? d1 : d0;
    const double su = gather16_edge<OP, 0>(tu, nf, NX, NY, k, slot, k < N);
    const double sv = gather16_edge<OP, 1>(tv, nf, NX, NY, k, slot, k < N);
    if (slot == 0 && k < N) {
        double r;
        if (mode == ILM_NORMAL) r = __dadd_rn(__dmul_rn(nx[k], su), __dmul_rn(ny[k], sv));
        else r = __dsub_rn(__dmul_rn(nx[k], sv), __dmul_rn(ny[k], su));
        dst[k] = __dmul_rn(coef, r / div);
    }
}
// op 0: grad of Nodes{Primal} (GLinvD), op 1: curl of Nodes{Dual} (CLinvCT)
int launch_sprobe_post(ilm_plan* p, int op, int mode, int ncol, const double* g0, const double* g1, double div, double coef,
                       double* d0, double* d1) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16;
    const dim3 grid((threads + 127) / 128, ncol);
    if (op == 0)
        k_sprobe_post<0><<<grid, 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]), p->g.NX, p->g.NY, g0, g1,
                                                      p->nx, p->ny, mode, div, coef, d0, d1);
    else
        k_sprobe_post<1><<<grid, 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]), p->g.NX, p->g.NY, g0, g1,
                                                      p->nx, p->ny, mode, div, coef, d0, d1);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// Fused pre/post kernels of the VectorData curl probes (create_CLinvCT / create_CL2invCT on a vector cache,
// src/matrix_operators.jl:40-61,102-125), two columns per launch:
//   pre : rows [rlo, rhi) of the Nodes{Dual} right-hand side = C^T R_f e_c / dx: the curl of ONE W x W edge
//         patch (u patch for c < N, v patch otherwise), zero elsewhere in the rows (other rows never read)
//   post: A[:, c] = coef * E_f (C s) / dx with the curl formed inside the gather (both edge components)
// The arithmetic is that of regularize! -> curl! and curl! -> interpolate! -> scale (same operations, same
// order); they replace 12 + 10 launches per column pair.
__global__ void k_vcurl_probe_pre(int N, int NX, int NY, TabView tu, TabView tv, const double* __restrict__ wRu,
                                  const double* __restrict__ wRv, int col0, int ncol, double* __restrict__ g0,
                                  double* __restrict__ g1, int rlo, int rhi, double div) {
    const int q = blockIdx.y;
    if (q >= ncol) return;
    const int c = col0 + q, comp = c < N ? 0 : 1, k = comp ? c - N : c;
    const TabView& t = comp ? tv : tu;
    const double* wR = comp ? wRv : wRu;
    double* out = q ? g1 : g0;
    const int ci = t.i0[k], cj = t.j0[k], W = t.W;
    auto patch = [&](int i, int j) -> double {              // R_f e_c at edge (i, j) of its component
        const int a = i - ci, b = j - cj;
        if (a < 0 || a >= W || b < 0 || b >= W || i < 0 || i >= t.mx || j < 0 || j >= t.my) return 0.0;
        return wR[(size_t)k * W * W + b * W + a];
    };
    const size_t n = (size_t)(rhi - rlo) * NX;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % NX), j = rlo + (int)(idx / NX);
        double val = 0.0;
        if (i >= 1 && i <= NX - 2 && j >= 1 && j <= NY - 2)
            val = (comp == 0 ? __dsub_rn(patch(i, j - 1), patch(i, j)) : __dsub_rn(patch(i, j), patch(i - 1, j))) / div;
        out[(size_t)j * NX + i] = val;
    }
}
int launch_vcurl_probe_pre(ilm_plan* p, int col0, int ncol, double* g0, double* g1, int rlo, int rhi, double div) {
    const size_t n = (size_t)(rhi - rlo) * p->g.NX;
    int bx = (int)((n + 255) / 256);
    if (bx > 2 * p->nsm) bx = 2 * p->nsm;
    if (bx < 1) bx = 1;
    k_vcurl_probe_pre<<<dim3(bx, ncol), 256, 0, p->stream>>>(p->N, p->g.NX, p->g.NY, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]),
                                                          p->tab[ILM_XEDGES].wR, p->tab[ILM_YEDGES].wR, col0, ncol, g0, g1, rlo, rhi, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

__global__ void k_vcurl_probe_post(int N, TabView tu, TabView tv, int NX, int NY, const double* __restrict__ g0,
                                   const double* __restrict__ g1, double div, double coef, double* __restrict__ d0,
                                   double* __restrict__ d1) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = gid >> 4, slot = gid & 15;
    const double* s = blockIdx.y ? g1 : g0;
    double* dst = blockIdx.y ? d1 : d0;
    const double su = gather16_edge<1, 0>(tu, s, NX, NY, k, slot, k < N);
    const double sv = gather16_edge<1, 1>(tv, s, NX, NY, k, slot, k < N);
    if (slot == 0 && k < N) {
        dst[k] = __dmul_rn(coef, su / div);
        dst[N + k] = __dmul_rn(coef, sv / div);
    }
}
int launch_vcurl_probe_post(ilm_plan* p, int ncol, const double* g0, const double* g1, double div, double coef, double* d0,
                            double* d1) {
    if (p->N == 0) return ILM_OK;
    const int threads = p->N * 16;
    k_vcurl_probe_post<<<dim3((threads + 127) / 128, ncol), 128, 0, p->stream>>>(p->N, view(p->tab[ILM_XEDGES]), view(p->tab[ILM_YEDGES]),
                                                                                  p->g.NX, p->g.NY, g0, g1, div, coef, d0, d1);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

// ------------------------------------------------------------------ mask products
// w *= grid_interpolate(mask -> layout of w), the body of _scalar_mask_product! / _vector_mask_product!
// (src/surface_operators.jl:880-923): product!(w, mask, w) on the mask's own layout, otherwise the
// mask is first averaged onto w's staggered positions (CartesianGrids grid_interpolate!: 2-point
// mean along a direction in which the two layouts are offset by half a cell, 4-point mean
// primal<->dual nodes).  Entries of w whose stencil leaves the mask field get the factor 0 (upstream
// leaves its zero-initialised scratch untouched there).  hx, hy = twice the layout shift.
__global__ void k_mask_product(double* __restrict__ w, int mx, int my, int hxw, int hyw, const double* __restrict__ m,
                               int mmx, int mmy, int hxm, int hym, int complementary) {
    const size_t n = (size_t)mx * my;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int dxo = hxm - hxw, dyo = hym - hyw;          // -1, 0, +1
    const int a0 = dxo < 0 ? -1 : 0, a1 = dxo > 0 ? 1 : 0;
    const int b0 = dyo < 0 ? -1 : 0, b1 = dyo > 0 ? 1 : 0;
    const double scale = 1.0 / (double)((a1 - a0 + 1) * (b1 - b0 + 1));
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % mx), j = (int)(idx / mx);          // 0-based
        double fac = 0.0;
        if (i + a0 >= 0 && i + a1 < mmx && j + b0 >= 0 && j + b1 < mmy) {
            double row[2] = {0.0, 0.0};
            for (int b = b0; b <= b1; ++b) {
                double acc = 0.0;
                for (int a = a0; a <= a1; ++a) {
                    double v = m[(size_t)(j + b) * mmx + (i + a)];
                    if (complementary) v = __dsub_rn(1.0, v);
                    acc = (a == a0) ? v : __dadd_rn(acc, v);
                }
                row[b - b0] = acc;
            }
            fac = __dmul_rn(scale, (b1 > b0) ? __dadd_rn(row[0], row[1]) : row[0]);
        }
        w[idx] = __dmul_rn(fac, w[idx]);
    }
}
int launch_mask_product(ilm_plan* p, double* w, int wlayout, const double* m, int mlayout, int complementary) {
    const LayoutInfo lw = layout_info(wlayout, p->g.NX, p->g.NY), lm = layout_info(mlayout, p->g.NX, p->g.NY);
    const size_t n = lw.n();
    if (n == 0) return ILM_OK;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)p->nsm * 16;
    if (blocks > cap) blocks = cap;
    k_mask_product<<<(unsigned)blocks, 256, 0, p->stream>>>(w, lw.mx, lw.my, (int)(2 * lw.sx), (int)(2 * lw.sy), m, lm.mx, lm.my,
                                                           (int)(2 * lm.sx), (int)(2 * lm.sy), complementary);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

}  // namespace ilm

// =====================================================================================
// Convective terms (src/grid_operators.jl:258-434): v.grad(p), v.grad(u), w x v.  The reference
// composes them from grad! / grid_interpolate! / product! / transpose! over three EdgeGradient-sized
// temporaries (7-9 sweeps); here each is ONE sweep that evaluates the intermediates on the fly from
// the (cached) neighbours, in the reference's operation order (no FMA contraction).  Intermediate
// entries whose stencil leaves the source field are 0, as in the zero-filled temporaries upstream.
// All indices 0-based; u: NX x (NY-1), v: (NX-1) x NY, primal: (NX-1) x (NY-1), dual: NX x NY.
// =====================================================================================
namespace ilm {

struct GridDims { int NX, NY; };
__device__ __forceinline__ double ld_u(const double* u, GridDims g, int a, int b) { return u[(size_t)b * g.NX + a]; }
__device__ __forceinline__ double ld_v(const double* v, GridDims g, int a, int b) { return v[(size_t)b * (g.NX - 1) + a]; }
__device__ __forceinline__ double ld_p(const double* p, GridDims g, int a, int b) { return p[(size_t)b * (g.NX - 1) + a]; }
__device__ __forceinline__ double ld_d(const double* d, GridDims g, int a, int b) { return d[(size_t)b * g.NX + a]; }
__device__ __forceinline__ double mean2(double x, double y) { return __dmul_rn(0.5, __dadd_rn(x, y)); }

// ---- v . grad p on Nodes{Primal} (:318-327) ---------------------------------------------------
__device__ __forceinline__ double cd_tu(const double* u, const double* p, GridDims g, int a, int b) {     // u * (grad p).u at x-edge (a,b)
    if (a < 1 || a > g.NX - 2) return 0.0;
    return __dmul_rn(ld_u(u, g, a, b), __dsub_rn(ld_p(p, g, a, b), ld_p(p, g, a - 1, b)));
}
__device__ __forceinline__ double cd_tv(const double* v, const double* p, GridDims g, int a, int b) {     // v * (grad p).v at y-edge (a,b)
    if (b < 1 || b > g.NY - 2) return 0.0;
    return __dmul_rn(ld_v(v, g, a, b), __dsub_rn(ld_p(p, g, a, b), ld_p(p, g, a, b - 1)));
}
__global__ void k_convective_scalar(GridDims g, const double* __restrict__ u, const double* __restrict__ v,
                                    const double* __restrict__ p, double* __restrict__ out, double div) {
    const int mx = g.NX - 1, my = g.NY - 1;
    const size_t n = (size_t)mx * my, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % mx), j = (int)(idx / mx);
        const double su = mean2(cd_tu(u, p, g, i, j), cd_tu(u, p, g, i + 1, j));
        const double sv = mean2(cd_tv(v, p, g, i, j), cd_tv(v, p, g, i, j + 1));
        out[idx] = __ddiv_rn(__dadd_rn(su, sv), div);
    }
}

// ---- w x v on Edges (:419-434) ---------------------------------------------------------------
__device__ __forceinline__ double wv_m1(const double* v, const double* w, GridDims g, int a, int b) {     // (-v at dual node) * w
    if (a < 1 || a > g.NX - 2) return 0.0;
    return __dmul_rn(__dmul_rn(-1.0, mean2(ld_v(v, g, a - 1, b), ld_v(v, g, a, b))), ld_d(w, g, a, b));
}
__device__ __forceinline__ double wv_m2(const double* u, const double* w, GridDims g, int a, int b) {     // (u at dual node) * w
    if (b < 1 || b > g.NY - 2) return 0.0;
    return __dmul_rn(mean2(ld_u(u, g, a, b - 1), ld_u(u, g, a, b)), ld_d(w, g, a, b));
}
__global__ void k_w_cross_v(GridDims g, const double* __restrict__ w, const double* __restrict__ u, const double* __restrict__ v,
                            double* __restrict__ ou, double* __restrict__ ov) {
    const size_t nu = (size_t)g.NX * (g.NY - 1), nv = (size_t)(g.NX - 1) * g.NY, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nu + nv; idx += stride) {
        if (idx < nu) {
            const int i = (int)(idx % g.NX), j = (int)(idx / g.NX);
            ou[idx] = mean2(wv_m1(v, w, g, i, j), wv_m1(v, w, g, i, j + 1));
        } else {
            const size_t k = idx - nu;
            const int i = (int)(k % (g.NX - 1)), j = (int)(k / (g.NX - 1));
            ov[k] = mean2(wv_m2(u, w, g, i, j), wv_m2(u, w, g, i + 1, j));
        }
    }
}

// ---- c . grad (u, v) on Edges (:329-375): c = advecting velocity, (u, v) = advected field ------
__device__ __forceinline__ double cv_p0(const double* cu, const double* u, GridDims g, int a, int b) {    // primal: cu@centre * dudx
    return __dmul_rn(mean2(ld_u(cu, g, a, b), ld_u(cu, g, a + 1, b)), __dsub_rn(ld_u(u, g, a + 1, b), ld_u(u, g, a, b)));
}
__device__ __forceinline__ double cv_p3(const double* cv, const double* v, GridDims g, int a, int b) {    // primal: cv@centre * dvdy
    return __dmul_rn(mean2(ld_v(cv, g, a, b), ld_v(cv, g, a, b + 1)), __dsub_rn(ld_v(v, g, a, b + 1), ld_v(v, g, a, b)));
}
__device__ __forceinline__ double cv_p1(const double* cv, const double* u, GridDims g, int a, int b) {    // dual: cv@node * dudy
    if (a < 1 || a > g.NX - 2 || b < 1 || b > g.NY - 2) return 0.0;     // cv@node needs a in range, dudy needs both
    return __dmul_rn(mean2(ld_v(cv, g, a - 1, b), ld_v(cv, g, a, b)), __dsub_rn(ld_u(u, g, a, b), ld_u(u, g, a, b - 1)));
}
__device__ __forceinline__ double cv_p2(const double* cu, const double* v, GridDims g, int a, int b) {    // dual: cu@node * dvdx
    if (a < 1 || a > g.NX - 2 || b < 1 || b > g.NY - 2) return 0.0;
    return __dmul_rn(mean2(ld_u(cu, g, a, b - 1), ld_u(cu, g, a, b)), __dsub_rn(ld_v(v, g, a, b), ld_v(v, g, a - 1, b)));
}
__global__ void k_convective_vector(GridDims g, const double* __restrict__ cu, const double* __restrict__ cv,
                                    const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ ou,
                                    double* __restrict__ ov, double div) {
    const size_t nu = (size_t)g.NX * (g.NY - 1), nv = (size_t)(g.NX - 1) * g.NY, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nu + nv; idx += stride) {
        if (idx < nu) {
            const int i = (int)(idx % g.NX), j = (int)(idx / g.NX);
            const double s0 = (i >= 1 && i <= g.NX - 2) ? mean2(cv_p0(cu, u, g, i - 1, j), cv_p0(cu, u, g, i, j)) : 0.0;
            const double s1 = mean2(cv_p1(cv, u, g, i, j), cv_p1(cv, u, g, i, j + 1));
            ou[idx] = __ddiv_rn(__dadd_rn(s0, s1), div);
        } else {
            const size_t k = idx - nu;
            const int i = (int)(k % (g.NX - 1)), j = (int)(k / (g.NX - 1));
            const double s2 = mean2(cv_p2(cu, v, g, i, j), cv_p2(cu, v, g, i + 1, j));
            const double s3 = (j >= 1 && j <= g.NY - 2) ? mean2(cv_p3(cv, v, g, i, j - 1), cv_p3(cv, v, g, i, j)) : 0.0;
            ov[k] = __ddiv_rn(__dadd_rn(s2, s3), div);
        }
    }
}

// ---- v . grad w on Nodes{Dual} (:272-288, 329-341) -----------------------------------------------
// Edges{Dual} temporaries: the x-component lives at the primal y-edge positions ((NX-1) x NY), the
// y-component at the primal x-edge positions (NX x (NY-1)).  grad!(Edges{Dual}, w), the 4-point means of the
// primal velocity at those positions, the product and the 2-point means back to the dual nodes.
__device__ __forceinline__ double cdw_tx(const double* u, const double* w, GridDims g, int a, int b) {    // at (a, b) of (NX-1) x NY
    if (b < 1 || b > g.NY - 2) return 0.0;                       // u (NX x (NY-1)) averaged over rows b-1, b and columns a, a+1
    const double r0 = __dadd_rn(ld_u(u, g, a, b - 1), ld_u(u, g, a + 1, b - 1));
    const double r1 = __dadd_rn(ld_u(u, g, a, b), ld_u(u, g, a + 1, b));
    const double cu = __dmul_rn(0.25, __dadd_rn(r0, r1));
    return __dmul_rn(cu, __dsub_rn(ld_d(w, g, a + 1, b), ld_d(w, g, a, b)));
}
__device__ __forceinline__ double cdw_ty(const double* v, const double* w, GridDims g, int a, int b) {    // at (a, b) of NX x (NY-1)
    if (a < 1 || a > g.NX - 2) return 0.0;                       // v ((NX-1) x NY) averaged over columns a-1, a and rows b, b+1
    const double r0 = __dadd_rn(ld_v(v, g, a - 1, b), ld_v(v, g, a, b));
    const double r1 = __dadd_rn(ld_v(v, g, a - 1, b + 1), ld_v(v, g, a, b + 1));
    const double cv = __dmul_rn(0.25, __dadd_rn(r0, r1));
    return __dmul_rn(cv, __dsub_rn(ld_d(w, g, a, b + 1), ld_d(w, g, a, b)));
}
__global__ void k_convective_dual(GridDims g, const double* __restrict__ u, const double* __restrict__ v,
                                  const double* __restrict__ w, double* __restrict__ out, double div) {
    const size_t n = (size_t)g.NX * g.NY, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx % g.NX), j = (int)(idx / g.NX);
        const double sx = (i >= 1 && i <= g.NX - 2) ? mean2(cdw_tx(u, w, g, i - 1, j), cdw_tx(u, w, g, i, j)) : 0.0;
        const double sy = (j >= 1 && j <= g.NY - 2) ? mean2(cdw_ty(v, w, g, i, j - 1), cdw_ty(v, w, g, i, j)) : 0.0;
        out[idx] = __ddiv_rn(__dadd_rn(sx, sy), div);
    }
}

static unsigned sweep_blocks(const ilm_plan* p, size_t n) {
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)p->nsm * 16;
    return (unsigned)(blocks > cap ? cap : (blocks ? blocks : 1));
}
int launch_convective_scalar(ilm_plan* p, const double* u, const double* v, const double* pn, double* out, double div) {
    const GridDims g{p->g.NX, p->g.NY};
    k_convective_scalar<<<sweep_blocks(p, (size_t)(g.NX - 1) * (g.NY - 1)), 256, 0, p->stream>>>(g, u, v, pn, out, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}
int launch_convective_dual(ilm_plan* p, const double* u, const double* v, const double* w, double* out, double div) {
    const GridDims g{p->g.NX, p->g.NY};
    k_convective_dual<<<sweep_blocks(p, (size_t)g.NX * g.NY), 256, 0, p->stream>>>(g, u, v, w, out, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}
int launch_w_cross_v(ilm_plan* p, const double* w, const double* u, const double* v, double* ou, double* ov) {
    const GridDims g{p->g.NX, p->g.NY};
    k_w_cross_v<<<sweep_blocks(p, (size_t)g.NX * (g.NY - 1) + (size_t)(g.NX - 1) * g.NY), 256, 0, p->stream>>>(g, w, u, v, ou, ov);
    ILM_LAUNCHED(p);
    return ILM_OK;
}
int launch_convective_vector(ilm_plan* p, const double* cu, const double* cv, const double* u, const double* v, double* ou,
                             double* ov, double div) {
    const GridDims g{p->g.NX, p->g.NY};
    k_convective_vector<<<sweep_blocks(p, (size_t)g.NX * (g.NY - 1) + (size_t)(g.NX - 1) * g.NY), 256, 0, p->stream>>>(g, cu, cv, u, v, ou, ov, div);
    ILM_LAUNCHED(p);
    return ILM_OK;
}

}  // namespace ilm


// ilm_band.cu -- "pass D": the column step of the convolution when the right-hand side has only a few
// non-zero rows (Schur probes: `R e_c`, `D_s e_c`, `C_s^T e_c` are WxW patches; create_RTLinvR and its
// siblings, src/matrix_operators.jl:9-215, call inverse_laplacian! on exactly such fields).
//
// With x_r(kx) the x-spectrum of non-zero row r (pass A) the y-direction needs no transform at all:
//     Y_n(kx) = sum_{r in [rlo, rhi)}  x_r(kx) * Gx(kx, |n - r|),      Gx(kx, d) = FFT_x of the mirrored kernel row d
// (real and even in kx).  That is <= NR real x complex FMAs per stored spectrum entry instead of a forward and an
// inverse length-2Ly transform, and the pass becomes a pure HBM stream: it writes the output rows of S2 (16 B per
// entry) and reads one real multiplier per entry.  Pass C then inverts along x as before.
//
// Thread = one x-frequency (px, m), walking down a run of output rows with the NR multipliers of the current row
// held in a register ring (one new coalesced load per row: Gx is stored transposed, [d][column], columns
// mirror-reduced like Ghat); x_r(kx) stays in registers.  Stores go to a ROW-MAJOR S2 ([row - olo][px][m],
// ConvArgs::s2_rowmajor): a warp writes 512 contiguous bytes per row and pass C fetches whole rows with one bulk copy
// each (first version: the 2x2-tile layout of the transform path, 32-byte pieces 131 KB apart -- 0.115 ms per launch).
//
// Patch mode (create_RTLinvR, PatchSrc): the non-zero rows are the two W x W DDF windows of the column pair, so the thread
// sums x_r(kx) = sum_q i^q sum_a wR_q[r][a] w^{kx (i0_q + a)} itself at kernel start -- no pre-operator, no pass A.  The row
// range [olo, ohi) is free: the symmetric Schur build (ilm_api.cu) asks for the rows from the pair's own windows upwards.
//
// The result is the same linear convolution as the transform path up to rounding (the sum over <= NR terms is
// exact arithmetic on the same x-spectrum); tests compare both with the oracle at 1e-12.
#include <cstdlib>

#include "ilm_internal.h"

namespace ilm {

// Gx table from the x-spectrum of h = e_i e_j (K - c0) (launch_lgf_prep; e = 1 at index 0, 2 elsewhere) that
// pass A has just left in S:  Re S_h[kx][d] = e_d * Gx(kx, d).
__global__ void k_gxt_extract(ConvGeom g, const double2* __restrict__ S, double* __restrict__ gxt, int ldc, int NY, double scale) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (t >= 2 * g.Lx || d >= NY) return;
    const int px = t / g.Lx, m = t - px * g.Lx;
    if (!ghat_is_rep(g, px, m)) return;
    const double v = S[s_index(g, px, m, d)].x * (d ? 0.5 : 1.0) * scale;
    gxt[(size_t)d * ldc + ghat_col(g, px, m)] = v;
}

// L2 policies: the multiplier rows (68 MB at 4096^2, the same rows for every column pair of a Schur build) are kept,
// the spectrum rows (268 MB per launch, read once by pass C) are streamed through
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_stream() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ld_keep(const double* ptr, unsigned long long pol) {
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol));      // read-only table: may be hoisted
    return v;
}
__device__ __forceinline__ void st_stream(double2* ptr, double re, double im, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(ptr), "d"(re), "d"(im), "l"(pol) : "memory");
}

template <int NR, int MINB>
__global__ void __launch_bounds__(256, MINB) k_passD(ConvGeom g, const double2* __restrict__ S, double2* __restrict__ S2,
                                               const double* __restrict__ gxt, int ldc, int dmax, int rlo, int nrows,
                                               int olo, int ohi, int chunk, PatchSrc ps) {
    const int t = blockIdx.x * 256 + threadIdx.x;            // x-frequency slot
    if (t >= 2 * g.Lx) return;
    const int px = t / g.Lx, m = t - px * g.Lx;
    const int n0 = olo + blockIdx.y * chunk;
    const int n1 = min(n0 + chunk, ohi);
    if (n0 >= n1) return;
    const unsigned long long keep = l2_policy_keep(), stream = l2_policy_stream();
    auto sidx = [&](int row) { return s_index(g, px, m, row); };
    double2 x[NR];
    if (ps.wR) {
        // patch mode: x_r(kx) = sum_q i^q w^{kx i0_q} sum_a wR_q[r - (j0_q - rlo)][a] w^{kx a},  w = exp(-2 pi i / PX):
        // the x-spectrum of the W x W windows summed directly (what pass A computes from the grid rows)
#pragma unroll
        for (int r = 0; r < NR; ++r) x[r] = cmk(0.0, 0.0);
        const int kx = 2 * m + px;
        const unsigned mask = (unsigned)(2 * g.Lx - 1);
        const double2 gw = ps.wl2x[kx];
        for (int q = 0; q < ps.ncol; ++q) {
            const int col = q ? ps.col1 : ps.col0;
            const int ci = ps.i0[col], cj = ps.j0[col];
            const double* w = ps.wR + (size_t)col * ps.W * ps.W;
            double2 pw = ps.wl2x[(unsigned)(kx * ci) & mask];          // w^{kx (i0 + a)}, a = 0
            if (q) pw = cmk(-pw.y, pw.x);                              // second column rides the imaginary part
            for (int a = 0; a < ps.W; ++a) {
                const int i = ci + a;
                if (i >= 0 && i < ps.mx) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const int b = rlo + r - cj;
                        if (r < nrows && b >= 0 && b < ps.W) {
                            const double wv = w[b * ps.W + a];
                            x[r].x = fma(wv, pw.x, x[r].x);
                            x[r].y = fma(wv, pw.y, x[r].y);
                        }
                    }
                }
                pw = cmul(pw, gw);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < NR; ++r) x[r] = r < nrows ? S[sidx(rlo + r)] : cmk(0.0, 0.0);
    }
    const double* gc = gxt + ghat_col(g, px, m);
    auto gload = [&](int e) {                                // multiplier of row distance |e| (clamped: rows past the
        int d = e < 0 ? -e : e;                              // run are computed but never stored)
        d = d < dmax ? d : dmax;
        return ld_keep(gc + (size_t)d * ldc, keep);
    };
    // offsets e = n - rlo; the multiplier of offset o lives in ring slot (o mod NR).  Start at the multiple of NR
    // at or below the first row so that slots are compile-time constants inside the unrolled body.
    const int e0 = n0 - rlo;
    const int es = e0 - (((e0 % NR) + NR) % NR);
    constexpr bool PIPE = NR <= 10;                          // wider windows keep the registers for x_r
    double ring[NR], nxt[PIPE ? NR : 1];
#pragma unroll
    for (int q = 1; q < NR; ++q) ring[NR - q] = gload(es - q);
    ring[0] = 0.0;
    if constexpr (PIPE) {
#pragma unroll
        for (int s = 0; s < NR; ++s) nxt[s] = gload(es + s);
    }
    const int eend = n1 - rlo;
    double2* out = S2 + ((size_t)(es + rlo - olo) * 2 + px) * (size_t)g.Lx + m;      // s2rm_index of row es + rlo
    const size_t rstride = (size_t)2 * g.Lx;
    for (int eb = es; eb < eend; eb += NR, out += NR * rstride) {
        double cur[NR];
        if constexpr (PIPE) {
#pragma unroll
            for (int s = 0; s < NR; ++s) cur[s] = nxt[s];
            if (eb + NR < eend) {                            // the next block's multipliers travel while this one is summed
#pragma unroll
                for (int s = 0; s < NR; ++s) nxt[s] = gload(eb + NR + s);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NR; ++s) cur[s] = gload(eb + s);
        }
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            const int n = eb + s + rlo;
            ring[s] = cur[s];
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double gg = ring[(s - r + NR) % NR];
                re = fma(x[r].x, gg, re);
                im = fma(x[r].y, gg, im);
            }
            if (n >= n0 && n < n1) st_stream(out + s * rstride, re, im, stream);
        }
    }
}

int conv_build_gxt(ilm_plan* p, const ConvArgs& a, ConvKernel& k, double factor) {
    const int NY = p->g.NY;
    k.gxt_ld = (p->Lx + 16) & ~15;
    k.gxt_rows = NY;
    ILM_CUDA(cudaMalloc(&k.gxt, (size_t)k.gxt_ld * NY * sizeof(double)));
    const double scale = 1.0 / (2.0 * (double)p->Lx * factor);
    const int tpb = 128;
    k_gxt_extract<<<dim3((2 * p->Lx + tpb - 1) / tpb, NY), tpb, 0, p->stream>>>(a.g, a.S, k.gxt, k.gxt_ld, NY, scale);
    ILM_CUDA(cudaGetLastError());
    p->launches += 1;
    return ILM_OK;
}

int conv_band_max_rows() { return 16; }

// S (rows [rlo, rhi) valid) -> S2 rows [olo, ohi)
int conv_launch_band(ilm_plan* p, const ConvArgs& a, const ConvKernel& k, const PatchSrc* psrc, cudaStream_t st) {
    const PatchSrc ps = psrc ? *psrc : PatchSrc{};
    if (!st) st = p->stream;
    const int nrows = a.rhi - a.rlo;
    const int nout = a.ohi - a.olo;
    if (nrows < 1 || nrows > conv_band_max_rows() || nout < 1) { set_error("conv_launch_band: bad row range"); return ILM_EINVAL; }
    const int gx = (2 * a.g.Lx + 255) / 256;
    // about 8 resident blocks per SM; runs of at least 32 rows so that the ring preload stays a small overhead
    int nchunks = (p->nsm * 8 + gx - 1) / gx;
    int chunk = (nout + nchunks - 1) / nchunks;
    if (chunk < 32) chunk = 32;
    chunk = (chunk + 1) & ~1;
    nchunks = (nout + chunk - 1) / chunk;
    const dim3 grid(gx, nchunks);
    const int dmax = k.gxt_rows - 1;
#define ILM_BAND(NR, MINB) k_passD<NR, MINB><<<grid, 256, 0, st>>>(a.g, a.S, a.S2, k.gxt, k.gxt_ld, dmax, a.rlo, nrows, a.olo, a.ohi, chunk, ps)
    static const int minb = getenv("ILM_BAND_MINB") ? atoi(getenv("ILM_BAND_MINB")) : 3;
    if (nrows <= 6) { if (minb >= 3) ILM_BAND(6, 3); else ILM_BAND(6, 2); }
    else if (nrows <= 8) { if (minb >= 3) ILM_BAND(8, 3); else ILM_BAND(8, 2); }
    else if (nrows <= 10) ILM_BAND(10, 2);
    else if (nrows <= 12) ILM_BAND(12, 2);
    else ILM_BAND(16, 2);
#undef ILM_BAND
    ILM_CUDA(cudaGetLastError());
    return ILM_OK;
}

}  // namespace ilm

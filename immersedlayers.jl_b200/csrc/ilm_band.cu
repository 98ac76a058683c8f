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
// mirror-reduced like Ghat); x_r(kx) stays in registers.  Stores go to the 2x2-tile spectrum layout of ilm_conv.cuh
// (lanes m, m+1 fill one 32-byte sector; consecutive rows complete the 64-byte tile).
//
// The result is the same linear convolution as the transform path up to rounding (the sum over <= NR terms is
// exact arithmetic on the same x-spectrum); tests compare both with the oracle at 1e-12.
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

template <int NR>
__global__ void __launch_bounds__(256) k_passD(ConvGeom g, const double2* __restrict__ S, double2* __restrict__ S2,
                                               const double* __restrict__ gxt, int ldc, int dmax, int rlo, int nrows,
                                               int olo, int ohi, int chunk) {
    const int t = blockIdx.x * 256 + threadIdx.x;            // x-frequency slot
    if (t >= 2 * g.Lx) return;
    const int px = t / g.Lx, m = t - px * g.Lx;
    const int n0 = olo + blockIdx.y * chunk;
    const int n1 = min(n0 + chunk, ohi);
    if (n0 >= n1) return;
    const size_t base = s_index(g, px, m, 0);
    auto sidx = [&](int row) { return base + ((size_t)(row >> 1) << 2) + (size_t)((row & 1) << 1); };
    double2 x[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) x[r] = r < nrows ? S[sidx(rlo + r)] : cmk(0.0, 0.0);
    const double* gc = gxt + ghat_col(g, px, m);
    auto gload = [&](int e) {                                // multiplier of row distance |e| (clamped: rows past the
        int d = e < 0 ? -e : e;                              // run are computed but never stored)
        d = d < dmax ? d : dmax;
        return __ldg(gc + (size_t)d * ldc);
    };
    // offsets e = n - rlo; the multiplier of offset o lives in ring slot (o mod NR).  Start at the multiple of NR
    // at or below the first row so that slots are compile-time constants inside the unrolled body.
    const int e0 = n0 - rlo;
    const int es = e0 - (((e0 % NR) + NR) % NR);
    double ring[NR];
#pragma unroll
    for (int q = 1; q < NR; ++q) ring[NR - q] = gload(es - q);
    ring[0] = 0.0;
    const int eend = n1 - rlo;
    for (int eb = es; eb < eend; eb += NR) {
#pragma unroll
        for (int s = 0; s < NR; ++s) {
            const int n = eb + s + rlo;
            ring[s] = gload(eb + s);
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double gg = ring[(s - r + NR) % NR];
                re = fma(x[r].x, gg, re);
                im = fma(x[r].y, gg, im);
            }
            if (n >= n0 && n < n1) S2[sidx(n)] = cmk(re, im);
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
int conv_launch_band(ilm_plan* p, const ConvArgs& a, const ConvKernel& k) {
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
#define ILM_BAND(NR) k_passD<NR><<<grid, 256, 0, p->stream>>>(a.g, a.S, a.S2, k.gxt, k.gxt_ld, dmax, a.rlo, nrows, a.olo, a.ohi, chunk)
    if (nrows <= 6) ILM_BAND(6);
    else if (nrows <= 8) ILM_BAND(8);
    else if (nrows <= 10) ILM_BAND(10);
    else if (nrows <= 12) ILM_BAND(12);
    else ILM_BAND(16);
#undef ILM_BAND
    ILM_CUDA(cudaGetLastError());
    return ILM_OK;
}

}  // namespace ilm

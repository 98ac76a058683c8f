// ilm_conv.cuh -- the three-pass pruned 2-D FFT convolution with a real,
// even kernel (lattice Green's function, integrating factor).
//
// Replaces CartesianGrids' CircularConvolution (zero-pad -> rfft -> multiply ->
// irfft -> crop) behind ImmersedLayers' inverse_laplacian!
// (src/grid_operators.jl:153-179) and Laplacian `L\w`.
//
// Two real fields w1, w2 are carried as one complex field z = w1 + i w2 (the
// kernel is real, so Re/Im never mix).  The padded transform size is
// PX x PY = 2Lx x 2Ly (powers of two >= 2NX-1, 2NY-1).  Because the inputs live
// in the first half of each padded direction and only the first half of the
// outputs is needed, every length-2L transform is pruned to two length-L
// transforms (even / odd output frequencies):
//     Z[2m]   = FFT_L(z)[m]            Z[2m+1] = FFT_L(z w_{2L}^n)[m]
//     z[n]    = IFFT_L(Z_even)[n] + conj(w_{2L}^n) IFFT_L(Z_odd)[n]
//   pass A : rows,    forward along x           field  -> S   (spectrum in x)
//   pass B : columns, forward y, * Ghat, inverse y     S -> S2
//   pass C : rows,    inverse along x           S2 -> field
// Ghat already contains 1/(PX PY), 1/L.factor and the far-field constant c0.
//
// Spectrum layout S[px][a][r2][rs][ms] (complex): x-frequency kx = 2m+px,
// m = 2a+ms; row = 2 r2 + rs.  A 2x2 tile is 64 contiguous bytes so that rows
// (passes A, C) and columns (pass B) both touch full 32-byte sectors.
#pragma once
#include "ilm_fft.cuh"

namespace ilm {

struct ConvGeom {
    int Lx, Ly;      // half padded lengths (PX = 2Lx, PY = 2Ly)
    int MY;          // number of field rows carried through the spectrum
    int MYp;         // MY rounded up to even
};

ILM_HD size_t s_index(const ConvGeom& g, int px, int m, int row) {
    return ((((size_t)px * (g.Lx >> 1) + (m >> 1)) * (size_t)(g.MYp >> 1) + (row >> 1)) << 2) +
           ((row & 1) << 1) + (m & 1);
}

ILM_HD size_t s_elems(const ConvGeom& g) { return (size_t)2 * g.Lx * g.MYp; }

// representative (mirror-reduced) x-frequency column of Ghat for (px, m)
ILM_HD int ghat_col(const ConvGeom& g, int px, int m) {
    if (px == 0) { int mm = g.Lx - m; return m < mm ? m : mm; }       // 0 .. Lx/2
    int mm = g.Lx - 1 - m;
    return (g.Lx >> 1) + 1 + (m < mm ? m : mm);                        // Lx/2+1 .. Lx
}
ILM_HD bool ghat_is_rep(const ConvGeom& g, int px, int m) {
    return px == 0 ? (m <= g.Lx - m) : (m <= g.Lx - 1 - m);
}
// Ghat layout: [col (Lx+1)][py (2)][k (Ly)] doubles
ILM_HD size_t ghat_elems(const ConvGeom& g) { return (size_t)(g.Lx + 1) * 2 * g.Ly; }

struct FieldRef {          // one real field: column-major, x fastest
    double* p;             // may be null (treated as zeros / not written)
    int mx, my;            // extents; leading dimension = mx
};

struct ConvArgs {
    ConvGeom g;
    FieldRef f1, f2;       // real / imaginary carrier
    double2* S;            // spectrum after pass A
    double2* S2;           // spectrum after pass B
    const double* Ghat;    // multiplier
    double* GhatOut;       // pass G output
    double gscale;         // pass G scale
    const double2* twx;    // twiddle table for Lx (global)
    const double2* twy;    // twiddle table for Ly (global)
};

// copy the twiddle table into shared memory (whole CTA, 512 threads)
template <int L, class Ctx>
ILM_HD void load_twiddles(Ctx& ctx, double2* tw_s, const double2* tw_g) {
    for (int i = ctx.grp * 256 + ctx.tid; i < FftCfg<L>::TW_TOTAL; i += 512) tw_s[i] = tw_g[i];
    ctx.sync_cta();
}

template <int L> ILM_HD double2 mod_fwd(const double2* tw, int j, int e) {
    return cmul(tw[FftCfg<L>::MOD_OFF + j], w32(e));
}

// ---------------------------------------------------------------- pass A
template <int L, class Ctx>
ILM_HD void passA_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    const int nwork = (a.g.MYp + 2 * F - 1) / (2 * F);
    for (int w = block; w < nwork; w += nblocks) {
        const int row = w * 2 * F + ctx.grp * F + f;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
        for (int px = 0; px < 2; ++px) {
            double2 v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                double re = (r1 && n < a.f1.mx) ? a.f1.p[(size_t)row * a.f1.mx + n] : 0.0;
                double im = (r2 && n < a.f2.mx) ? a.f2.p[(size_t)row * a.f2.mx + n] : 0.0;
                v[e] = cmk(re, im);
                if (px) v[e] = cmul(v[e], mod_fwd<L>(tw, j, e));
            }
            fft_regs<L, false>(v, ctx, xb, tw, j);
            if (row < a.g.MYp) {
#pragma unroll
                for (int e = 0; e < 16; ++e) a.S[s_index(a.g, px, j + e * T, row)] = v[e];
            }
        }
    }
}

// ---------------------------------------------------------------- pass B / G
// MODE 0: convolution step (S -> S2); MODE 1: build Ghat from Re(S)
template <int L, int MODE, class Ctx>
ILM_HD void passB_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twy);
    const int f = ctx.tid / T, j = ctx.tid % T;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    const int ntiles = a.g.Lx;                     // (Lx/2 tiles) x (2 parities)
    const int nwork = (ntiles + F - 1) / F;
    for (int w = block; w < nwork; w += nblocks) {
        const int tile = w * F + f;
        const bool live = tile < ntiles;
        const int px = live ? tile / (a.g.Lx >> 1) : 0;
        const int m = live ? ((tile % (a.g.Lx >> 1)) << 1) + ctx.grp : 0;
        const size_t gbase = (size_t)ghat_col(a.g, px, m) * 2 * a.g.Ly;
        const bool rep = MODE == 1 && live && ghat_is_rep(a.g, px, m);
        for (int py = 0; py < 2; ++py) {
            double2 v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                v[e] = (live && n < a.g.MYp) ? a.S[s_index(a.g, px, m, n)] : cmk(0.0, 0.0);
                if (MODE == 1) v[e].y = 0.0;
                if (py) v[e] = cmul(v[e], mod_fwd<L>(tw, j, e));
            }
            fft_regs<L, false>(v, ctx, xb, tw, j);
            if (MODE == 1) {
                if (rep) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        a.GhatOut[gbase + (size_t)py * a.g.Ly + j + e * T] = v[e].x * a.gscale;
                }
                continue;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const double gh = live ? a.Ghat[gbase + (size_t)py * a.g.Ly + j + e * T] : 0.0;
                v[e] = cmk(v[e].x * gh, v[e].y * gh);
            }
            fft_regs<L, true>(v, ctx, xb, tw, j);
            if (live) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int n = j + e * T;
                    if (n < a.g.MYp) {
                        const size_t idx = s_index(a.g, px, m, n);
                        if (py == 0) a.S2[idx] = v[e];
                        else a.S2[idx] = cadd(a.S2[idx], cmulc(v[e], mod_fwd<L>(tw, j, e)));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- pass C
template <int L, class Ctx>
ILM_HD void passC_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    const int nwork = (a.g.MYp + 2 * F - 1) / (2 * F);
    for (int w = block; w < nwork; w += nblocks) {
        const int row = w * 2 * F + ctx.grp * F + f;
        const bool live = row < a.g.MYp;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
        for (int px = 0; px < 2; ++px) {
            double2 v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e)
                v[e] = live ? a.S2[s_index(a.g, px, j + e * T, row)] : cmk(0.0, 0.0);
            fft_regs<L, true>(v, ctx, xb, tw, j);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                double2 y = px ? cmulc(v[e], mod_fwd<L>(tw, j, e)) : v[e];
                if (r1 && n < a.f1.mx) {
                    double* q = a.f1.p + (size_t)row * a.f1.mx + n;
                    *q = px ? *q + y.x : y.x;
                }
                if (r2 && n < a.f2.mx) {
                    double* q = a.f2.p + (size_t)row * a.f2.mx + n;
                    *q = px ? *q + y.y : y.y;
                }
            }
        }
    }
}

}  // namespace ilm

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
// A CTA is two 256-thread groups: group 0 owns the even half-transform and
// group 1 the odd one of the SAME rows/columns, so a row or column is read
// once, and the two halves of an inverse transform are combined through shared
// memory (no global round trip).  The next work item is prefetched into L2
// while the current one is transformed.
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

// L2 prefetch of `bytes` starting at `base`, spread over the CTA's 512 threads
template <class Ctx> ILM_HD void prefetch_range(Ctx& ctx, const void* base, size_t bytes) {
    const char* b = (const char*)base;
    for (size_t off = (size_t)(ctx.grp * 256 + ctx.tid) * 128; off < bytes; off += 512 * 128) ctx.prefetch_l2(b + off);
}

// ---------------------------------------------------------------- pass A
// work item = F rows; group g computes output parity px = g of those rows
template <int L, class Ctx>
ILM_HD void passA_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T, px = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    const int nwork = (a.g.MYp + F - 1) / F;
    for (int w = block; w < nwork; w += nblocks) {
        const int row = w * F + f;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
        const int wn = w + nblocks;                       // next work item of this CTA
        if (wn < nwork) {
            if (a.f1.p && wn * F < a.f1.my) prefetch_range(ctx, a.f1.p + (size_t)wn * F * a.f1.mx, (size_t)F * a.f1.mx * 8);
            if (a.f2.p && wn * F < a.f2.my) prefetch_range(ctx, a.f2.p + (size_t)wn * F * a.f2.mx, (size_t)F * a.f2.mx * 8);
        }
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

// ---------------------------------------------------------------- pass B / G
// work item = CPW consecutive x-frequency columns (tile-column order); group g
// owns y-parity py = g.  MODE 0: convolution (S -> S2); MODE 1: Ghat = Re(FFT_y(Re S)).
template <int L, int MODE, class Ctx>
ILM_HD void passB_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    constexpr int SUB = (F == 1) ? 2 : 1;        // a 2-column tile is never split between CTAs
    constexpr int CPW = F * SUB;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twy);
    const int f = ctx.tid / T, j = ctx.tid % T, py = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    double2* xb_odd = smem + C::GROUP_XBUF + f * C::XBUF;      // group 1's buffer of the same FFT slot
    const int ncols = 2 * a.g.Lx;
    const int nwork = (ncols + CPW - 1) / CPW;
    const size_t col_elems = (size_t)a.g.MYp;                   // complex elements per column
    for (int w = block; w < nwork; w += nblocks) {
        const int wn = w + nblocks;
        if (wn < nwork) {
            const size_t e0 = (size_t)wn * CPW * col_elems;
            size_t ne = (size_t)CPW * col_elems;
            if (e0 + ne > s_elems(a.g)) ne = s_elems(a.g) - e0;
            prefetch_range(ctx, a.S + e0, ne * sizeof(double2));
        }
#pragma unroll 1
        for (int sub = 0; sub < SUB; ++sub) {
            const int c = w * CPW + sub * F + f;
            const bool live = c < ncols;
            const int px = live ? c / a.g.Lx : 0;
            const int m = live ? c % a.g.Lx : 0;
            const size_t gbase = ((size_t)ghat_col(a.g, px, m) * 2 + py) * a.g.Ly;
            if (MODE == 0 && live) ctx.prefetch_l2(a.Ghat + gbase + (size_t)j * 16);    // T lines of 16 doubles
            double2 v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                v[e] = (live && n < a.g.MYp) ? a.S[s_index(a.g, px, m, n)] : cmk(0.0, 0.0);
                if (MODE == 1) v[e].y = 0.0;
                if (py) v[e] = cmul(v[e], mod_fwd<L>(tw, j, e));
            }
            fft_regs<L, false>(v, ctx, xb, tw, j);
            if constexpr (MODE == 1) {
                if (live && ghat_is_rep(a.g, px, m)) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) a.GhatOut[gbase + j + e * T] = v[e].x * a.gscale;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const double gh = live ? a.Ghat[gbase + j + e * T] : 0.0;
                    v[e] = cmk(v[e].x * gh, v[e].y * gh);
                }
                fft_regs<L, true>(v, ctx, xb, tw, j);
                // combine the two half transforms: y[n] = E[n] + conj(w^n) O[n]
                if (py) {
                    ctx.sync();                               // group 1 finished reading its exchange buffer
#pragma unroll
                    for (int e = 0; e < 16; ++e) xb[xpad(j + e * T)] = cmulc(v[e], mod_fwd<L>(tw, j, e));
                }
                ctx.sync_cta();
                if (!py && live) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = j + e * T;
                        if (n < a.g.MYp) a.S2[s_index(a.g, px, m, n)] = cadd(v[e], xb_odd[xpad(n)]);
                    }
                }
                ctx.sync_cta();                               // group 1 may reuse its buffer
            }
        }
    }
}

// ---------------------------------------------------------------- pass C
// work item = F rows; group g inverts parity px = g; group 0 combines and stores
template <int L, class Ctx>
ILM_HD void passC_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + 2 * C::GROUP_XBUF;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T, px = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    double2* xb_odd = smem + C::GROUP_XBUF + f * C::XBUF;
    const int nwork = (a.g.MYp + F - 1) / F;
    for (int w = block; w < nwork; w += nblocks) {
        const int row = w * F + f;
        const bool live = row < a.g.MYp;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
        double2 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = live ? a.S2[s_index(a.g, px, j + e * T, row)] : cmk(0.0, 0.0);
        // next row of this CTA: the 32-byte sectors this thread will read
        const int rown = row + nblocks * F;
        if (rown < a.g.MYp) {
#pragma unroll
            for (int e = 0; e < 16; ++e) ctx.prefetch_l2(&a.S2[s_index(a.g, px, j + e * T, rown)]);
        }
        fft_regs<L, true>(v, ctx, xb, tw, j);
        if (px) {
            ctx.sync();
#pragma unroll
            for (int e = 0; e < 16; ++e) xb[xpad(j + e * T)] = cmulc(v[e], mod_fwd<L>(tw, j, e));
        }
        ctx.sync_cta();
        if (!px) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                const double2 y = cadd(v[e], xb_odd[xpad(n)]);
                if (r1 && n < a.f1.mx) a.f1.p[(size_t)row * a.f1.mx + n] = y.x;
                if (r2 && n < a.f2.mx) a.f2.p[(size_t)row * a.f2.mx + n] = y.y;
            }
        }
        ctx.sync_cta();
    }
}

}  // namespace ilm

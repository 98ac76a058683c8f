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
    int MYp;         // MY rounded up to even (to a multiple of 2 rq)
    int rq;          // row interleave (0 / 1 = none): the rows are stored class by class, row n at position
                     // (n % rq) * (MYp / rq) + n / rq, so that the decimated sub-sequences n1 + Q n2 of the big column
                     // pass (ilm_conv_big.cuh, rq = Q) are contiguous runs of a column instead of every Q-th row
};

ILM_HD int s_row(const ConvGeom& g, int row) { return g.rq > 1 ? (row % g.rq) * (g.MYp / g.rq) + row / g.rq : row; }

ILM_HD size_t s_index(const ConvGeom& g, int px, int m, int row) {
    row = s_row(g, row);
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

// Fused interpolation of the Schur probes (create_RTLinvR: the post-operator is E on one table): pass C hands every
// inverted row to the points whose windows cross it instead of storing it.  part[k*W + b] = sum_a wE[k][b][a] y[row][i0_k + a]
// (complex: two probe columns); a small kernel then adds the W row sums of each point.
struct ProbeGather {
    const int* rowptr;     // my + 1 offsets into rowent
    const int* rowent;     // k*W + b of every in-field window row, sorted by row, then point
    const int* i0;         // first window column of each point
    const double* w;       // interpolation weights [k][b][a]
    int W, mx, my;
    double2* part;         // null = store the rows (every other caller)
    const unsigned* need;         // my x (Lx/16) column masks: bit e of need[row*T + j] = column j + e*T is read by a window
                                  // (null = every column); threads / warps without a needed column skip the last pass
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
    int skew_ns;           // start-up delay of one group (de-phases FP64 and shared-memory phases)
    const double2* wl2y;   // exp(-2 pi i n / (2 Ly)), n < 2 Ly (global; sparse forward transform of pass B)
    int rlo, rhi;          // input rows outside [rlo, rhi) are known to be zero in both fields
                           // (Schur probes: a 4x4 patch); pass A skips them, pass B reads zeros
    int olo, ohi;          // only output rows [olo, ohi) are needed (Schur probes: the rows under the
                           // interpolation windows); pass B stores and pass C inverts only those rows,
                           // the other rows of the output fields are left untouched.  olo is even.
    const double2* wl2x;   // exp(-2 pi i n / (2 Lx)), n < 2 Lx (only for Lx > 4096, ilm_conv_big.cuh)
    double2* scratch;      // per-CTA hand-off lines of the big column pass (2 groups x Ly complex per CTA)
    int wlo, whi;          // pass B visits the work items [wlo, whi) only (slab decomposition: the x-frequency
                           // columns this GPU owns); whi <= 0 = all of them
    ProbeGather eg;        // fused interpolation in pass C (Schur probes)
    double2* bigA;         // hand-off block of the split big column pass (ilm_conv_big.cuh): [column - bc0][p][n1][kappa]
    int bc0, bnc;          // ... for the columns [bc0, bc0 + bnc)
    int s2_rowmajor;       // S2 holds the output rows [olo, ohi) row by row, [row - olo][px][m] (written by the band
                           // pass of the Schur probes, ilm_band.cu): both that pass and pass C then stream whole rows
};

// row-major S2 of the probe path: one x-spectrum row is 2 Lx contiguous complex numbers (even frequencies, then odd)
ILM_HD size_t s2rm_index(const ConvArgs& a, int px, int m, int row) {
    return ((size_t)(row - a.olo) * 2 + px) * (size_t)a.g.Lx + m;
}

// named barriers 1,2 are the per-group barriers (Ctx::sync); these two carry the
// producer (odd half, group 1) -> consumer (even half, group 0) hand-off
constexpr int BAR_READY = 3, BAR_FREE = 4;

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

// spectrum index of element e of thread j when the transform runs along x (rows):
// m = j + e*T at fixed row; constant stride T*MYp for even T
template <int T> ILM_HD size_t row_elem(const ConvGeom& g, int px, int row, int j, int e, size_t i0) {
    if constexpr (T >= 2) return i0 + (size_t)e * T * (size_t)g.MYp;
    else return s_index(g, px, j + e * T, row);
}
// ... and along y (columns): n = j + e*T at fixed m; constant stride 2T for even T
template <int T> ILM_HD size_t col_elem(const ConvGeom& g, int px, int m, int j, int e, size_t i0) {
    if constexpr (T >= 2) return i0 + (size_t)e * (2 * T);
    else return s_index(g, px, m, j + e * T);
}

// ---------------------------------------------------------------- pass A
// work item = F rows; group g computes output parity px = g of those rows
template <int L, class Ctx>
ILM_HD void passA_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T, px = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    // work item = F rows (measured: pairing consecutive rows per CTA is slower here)
    constexpr int SUB = 1, RPW = F * SUB;
    const int nwork = (a.rhi - a.rlo + RPW - 1) / RPW;
    if (px) ctx.delay(a.skew_ns);
    for (int w = block; w < nwork; w += nblocks) {
        const int wn = w + nblocks;                       // next work item of this CTA
        if (wn < nwork) {
            const int rn = a.rlo + wn * RPW;
            if (a.f1.p && rn < a.f1.my) prefetch_range(ctx, a.f1.p + (size_t)rn * a.f1.mx, (size_t)RPW * a.f1.mx * 8);
            if (a.f2.p && rn < a.f2.my) prefetch_range(ctx, a.f2.p + (size_t)rn * a.f2.mx, (size_t)RPW * a.f2.mx * 8);
        }
#pragma unroll 1
      for (int sub = 0; sub < SUB; ++sub) {
        const int row = a.rlo + w * RPW + sub * F + f;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
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
        if (row < a.rhi) {
            const size_t i0 = s_index(a.g, px, j, row);
#pragma unroll
            for (int e = 0; e < 16; ++e) a.S[row_elem<T>(a.g, px, row, j, e, i0)] = v[e];
        }
      }
    }
}

// Forward half transform of a column whose only non-zero rows are [rlo, rhi) (Schur probe:
// a WxW patch): X_py[k] = sum_r x_r w_{2L}^{r (2k + py)} summed directly, k = j + e*T.
// About 300 FP64 instructions and no shared-memory exchange, against ~780 and two exchanges
// for the full transform; used when the row range is short.
constexpr int SPARSE_MAX_ROWS = 16;
template <int L, class Ctx>
ILM_HD void sparse_forward(double2* v, Ctx& ctx, const ConvArgs& a, double2* xb, int px, int m, int py, int j, bool live) {
    const int nrows = a.rhi - a.rlo;
    // both loads of this transform are issued up front: the column's non-zero rows (staged in the
    // exchange buffer for the whole FFT slot) and the twiddle of the first row; the twiddles of the
    // following rows come from the recurrence b_{r+1} = b_r * w_{2L}^{2j+py}
    const unsigned mask = (unsigned)(2 * L - 1);
    const double2 g = a.wl2y[(unsigned)(2 * j + py) & mask];
    double2 b = a.wl2y[((unsigned)a.rlo * (unsigned)(2 * j + py)) & mask];
    ctx.sync();                                            // the exchange buffer is free
    for (int r = j; r < nrows; r += FftCfg<L>::T) xb[r] = live ? a.S[s_index(a.g, px, m, a.rlo + r)] : cmk(0.0, 0.0);
    ctx.sync();
    // v[e] = sum_r t_r w_16^{(rlo + r) e} with t_r = x_r b_r: one 16-point DFT of the rows
    // (placed at 0..nrows-1) followed by the rotation w_16^{rlo e}
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (r < SPARSE_MAX_ROWS && r < nrows) {
            v[r] = cmul(xb[r], b);
            b = cmul(b, g);
        } else {
            v[r] = cmk(0.0, 0.0);
        }
    }
    fft16<false>(v);
    const int rl = a.rlo & 15;
#pragma unroll
    for (int e = 1; e < 16; ++e) v[e] = cmul(v[e], w16((rl * e) & 15));
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
    double2* tw = smem + C::TW_BASE;
    load_twiddles<L>(ctx, tw, a.twy);
    const int f = ctx.tid / T, j = ctx.tid % T, py = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF + f * C::XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF + f * L;           // odd -> even hand-off of this FFT slot
    const int ncols = 2 * a.g.Lx;
    const int nwork = (ncols + CPW - 1) / CPW;
    const size_t col_elems = (size_t)a.g.MYp;                   // complex elements per column
    if (MODE == 0) {
        if (!py) { ctx.arrive(BAR_FREE); ctx.delay(a.skew_ns); }
    }
    // Work order: inside each parity block the tile columns are visited from both ends
    // (0, last, 1, last-1, ...) so that a column and its mirror kx <-> PX-kx, which share one
    // Ghat column, are processed at about the same time by neighbouring CTAs: Ghat is read from
    // DRAM once and hit in L2 the second time (F == 1 only; smaller transforms keep the identity).
    const bool ranged = a.whi > 0;                              // a slab owns a contiguous column range: identity order
    auto tile_of = [&](int w) {
        if (F != 1 || ranged) return w;
        const int half = a.g.Lx >> 1, blk = w / half, t = w % half;
        return blk * half + ((t & 1) ? half - 1 - (t >> 1) : (t >> 1));
    };
    const int wbeg = ranged ? a.wlo : 0, wend = ranged ? (a.whi < nwork ? a.whi : nwork) : nwork;
    for (int w0 = wbeg + block; w0 < wend; w0 += nblocks) {
        const int w = tile_of(w0);
        const int wn0 = w0 + nblocks;
        if (wn0 < wend && a.rhi - a.rlo == a.g.MYp) {
            const size_t e0 = (size_t)tile_of(wn0) * CPW * col_elems;
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
            const size_t i0 = s_index(a.g, px, m, j);
            if (MODE == 0 && live) ctx.prefetch_l2(a.Ghat + gbase + (size_t)j * 16);    // T lines of 16 doubles
            double2 v[16];
            if (MODE == 0 && a.rhi - a.rlo <= SPARSE_MAX_ROWS) {
                sparse_forward<L>(v, ctx, a, xb, px, m, py, j, live);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int n = j + e * T;
                    v[e] = (live && n >= a.rlo && n < a.rhi) ? a.S[col_elem<T>(a.g, px, m, j, e, i0)] : cmk(0.0, 0.0);
                    if (MODE == 1) v[e].y = 0.0;
                    if (py) v[e] = cmul(v[e], mod_fwd<L>(tw, j, e));
                }
                fft_regs<L, false>(v, ctx, xb, tw, j);
            }
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
                    ctx.wait(BAR_FREE);                       // consumer has drained the previous column
#pragma unroll
                    for (int e = 0; e < 16; ++e) comb[j + e * T] = cmulc(v[e], mod_fwd<L>(tw, j, e));
                    ctx.arrive(BAR_READY);
                } else {
                    ctx.wait(BAR_READY);
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int n = j + e * T;
                            if (n >= a.olo && n < a.ohi) a.S2[col_elem<T>(a.g, px, m, j, e, i0)] = cadd(v[e], comb[n]);
                        }
                    }
                    ctx.arrive(BAR_FREE);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- pass C
// work item = RPW rows; group g inverts parity px = g; group 0 combines and stores.
// For L >= 512 the spectrum rows are fetched with bulk-tensor (TMA) copies straight
// into the group's exchange buffer, one step ahead: the copy of step s+1 is issued
// as soon as step s has read the buffer for the last time, and lands while the last
// butterfly pass, the combine and the stores of step s run.  The 2x2-tile gather
// (32-byte pieces) is done by the copy engine instead of 16 scattered loads per thread.
template <int L, class Ctx>
ILM_HD void passC_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<L>;
    constexpr int T = C::T, F = C::F;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<L>(ctx, tw, a.twx);
    const int f = ctx.tid / T, j = ctx.tid % T, px = ctx.grp;
    double2* xgrp = smem + ctx.grp * C::GROUP_XBUF;
    double2* xb = xgrp + f * C::XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF + f * L;
    double2* mbar = smem + C::MBAR_OFF + ctx.grp;
    // RPW consecutive rows per work item: the second row of each 2x2 tile is an
    // L2 hit (its sector came in with the first row's 64-byte DRAM access)
    constexpr int SUB = (F == 1) ? 2 : 1, RPW = F * SUB;
    const int nwork = (a.ohi - a.olo + RPW - 1) / RPW;
    const int nsteps = (block < nwork) ? ((nwork - block + nblocks - 1) / nblocks) * SUB : 0;
    // rows of step s for FFT slot ff
    auto step_row0 = [&](int s) { return a.olo + (block + (s / SUB) * nblocks) * RPW + (s % SUB) * F; };
    if constexpr (C::USE_TMA) {
        ctx.tma_init(mbar);
        if (nsteps > 0) ctx.tma_load_rows(a, px, step_row0(0), F, L, C::XBUF, xgrp, mbar);
    }
    if (!px) { ctx.arrive(BAR_FREE); ctx.delay(a.skew_ns); }
    for (int s = 0; s < nsteps; ++s) {
        const int row = step_row0(s) + f;
        const bool live = row < a.ohi;
        const bool r1 = live && a.f1.p && row < a.f1.my, r2 = live && a.f2.p && row < a.f2.my;
        double2 v[16];
        if constexpr (C::USE_TMA) {
            ctx.tma_wait(mbar, s & 1);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = xb[j + e * T];
        } else {
            const size_t i0 = s_index(a.g, px, j, row);
#pragma unroll
            for (int e = 0; e < 16; ++e)
                v[e] = live ? a.S2[a.s2_rowmajor ? s2rm_index(a, px, j + e * T, row) : row_elem<T>(a.g, px, row, j, e, i0)] : cmk(0.0, 0.0);
        }
        fft_head<L, true>(v, ctx, xb, tw, j);
        if constexpr (C::USE_TMA) {
            if (s + 1 < nsteps) {
                ctx.sync();                                   // every thread has left the exchange buffer
                ctx.tma_load_rows(a, px, step_row0(s + 1), F, L, C::XBUF, xgrp, mbar);
            }
        }
        // probe path with a column mask: only the columns under the interpolation windows are formed.  A warp none of
        // whose threads owns such a column skips the last butterfly pass (a third of the transform's FP64 work), and the
        // odd -> even hand-off touches the needed entries only.
        unsigned need = 0xffffu;
        bool skip_last = false;
        if (a.eg.part && a.eg.need) {
            need = (live && row < a.eg.my) ? a.eg.need[(size_t)row * T + j] : 0u;
            skip_last = !ctx.any(need != 0u);
        }
        if (!skip_last) fft_last_pass<L, true>(v, tw, j);
        if (px) {
            ctx.wait(BAR_FREE);
            if (need == 0xffffu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) comb[j + e * T] = cmulc(v[e], mod_fwd<L>(tw, j, e));
            } else if (need) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if ((need >> e) & 1u) comb[j + e * T] = cmulc(v[e], mod_fwd<L>(tw, j, e));
            }
            ctx.arrive(BAR_READY);
        } else if (a.eg.part) {
            // probe path: the finished row stays in the combine buffer and is gathered by the points whose
            // interpolation windows cross it (same term order inside a window row as the oracle: a ascending)
            ctx.wait(BAR_READY);
            if (need == 0xffffu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) { const int n = j + e * T; comb[n] = cadd(v[e], comb[n]); }
            } else if (need) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if ((need >> e) & 1u) { const int n = j + e * T; comb[n] = cadd(v[e], comb[n]); }
            }
            ctx.sync();
            if (live && row < a.eg.my) {
                const int W = a.eg.W;
                for (int q = a.eg.rowptr[row] + j; q < a.eg.rowptr[row + 1]; q += T) {
                    const int ent = a.eg.rowent[q];
                    const int i0 = a.eg.i0[ent / W];
                    const double* w = a.eg.w + (size_t)ent * W;
                    double re = 0.0, im = 0.0;
                    for (int c = 0; c < W; ++c) {
                        const int i = i0 + c;
                        if (i >= 0 && i < a.eg.mx) { re += w[c] * comb[i].x; im += w[c] * comb[i].y; }
                    }
                    a.eg.part[ent] = cmk(re, im);
                }
            }
            ctx.arrive(BAR_FREE);
        } else {
            ctx.wait(BAR_READY);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = j + e * T;
                const double2 y = cadd(v[e], comb[n]);
                if (r1 && n < a.f1.mx) a.f1.p[(size_t)row * a.f1.mx + n] = y.x;
                if (r2 && n < a.f2.mx) a.f2.p[(size_t)row * a.f2.mx + n] = y.y;
            }
            ctx.arrive(BAR_FREE);
        }
    }
}

}  // namespace ilm

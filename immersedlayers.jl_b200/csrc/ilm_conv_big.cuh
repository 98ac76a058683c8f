// ilm_conv_big.cuh -- the three convolution passes for half padded lengths L = Q * 4096
// (Q = 2, 4: grids of 4097..16384 cells per direction; BASELINE config C5b is 16384^2).
//
// A length-L transform does not fit one SM (16384 complex = 256 KB > 227 KB of shared memory), so
// the zero-padded length-2L transform is split by one radix-2Q decimation-in-frequency step on
// top of the 4096-point register/shared-memory transform of ilm_fft.cuh (M = 4096):
//
//   forward   Z[c + 2Q k2] = FFT_M( v_c )[k2],      v_c[n2] = sum_{n1<Q} z[n2 + M n1] w_2L^{(n2 + M n1) c}
//             for the 2Q residue classes c = p + 2 k1 (p = output parity handled by group p);
//             the even/odd pruning of ilm_conv.cuh is the Q = 1 case of the same formula.
//   inverse   z[n1 + Q n2] = Y_0[n2] + conj(w_2M^{n2}) Y_1[n2],   Y_p = IFFT_M( U_{n1}[2 kappa + p] ),
//             U_{n1}[k2] = conj(w_2L^{n1 k2}) sum_{k1<Q} Z[k2 + 2M k1] exp(+2 pi i n1 k1 / Q)
//             for the Q output residues n1; the two halves are combined exactly as in ilm_conv.cuh.
//
// In the column pass the radix-Q combination that prepares the inverse is thread-local: the 16
// registers of a thread after the forward class c hold Z[c + 2Q (j + 256 e)], and e = e' + (16/Q) k1
// are precisely the Q entries k2 + 2M k1 of one U.  The forward classes of a column therefore hand
// their results to the inverse residues through a per-CTA scratch line in global memory (2L complex
// per group, L2-resident), the only intermediate that exceeds shared memory.
//
// Pass A re-reads a row Q times (real input, L2 hits); pass C reads each spectrum entry Q times
// (once per output residue; the row stays in L2 between the residues of a work group).
#pragma once
#include "ilm_conv.cuh"

namespace ilm {

constexpr int BIG_M = 4096;

// z * exp(+2 pi i t / Q), Q = 2 or 4
template <int Q> ILM_HD double2 rotq(double2 z, int t) {
    if constexpr (Q == 2) return (t & 1) ? cmk(-z.x, -z.y) : z;
    else {
        switch (t & 3) {
        case 0: return z;
        case 1: return cmk(-z.y, z.x);
        case 2: return cmk(-z.x, -z.y);
        default: return cmk(z.y, -z.x);
        }
    }
}

// ---------------------------------------------------------------- pass A (rows, forward)
// work item = row; its 2Q classes c = px + 2 k1 are visited as Q steps (px, k1 pair): group g takes
// k1 = 2 * pair + g, so that the two groups write the two 16-byte halves (m even / odd) of the same
// 32-byte spectrum sectors at the same time and L2 merges them into full-sector DRAM writes.
template <int Q, class Ctx>
ILM_HD void passA_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twx);
    const int j = ctx.tid;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    const unsigned mask = 2u * (unsigned)a.g.Lx - 1u;
    const int nrows = a.rhi - a.rlo;
    for (int w = block; w < nrows; w += nblocks) {
        const int row = a.rlo + w;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
#pragma unroll 1
        for (int step = 0; step < Q; ++step) {
            const int px = step / (Q / 2), k1 = 2 * (step % (Q / 2)) + ctx.grp;
            const unsigned cls = (unsigned)(px + 2 * k1);
            double2 v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n2 = j + e * T;
                double2 acc = cmk(0.0, 0.0);
#pragma unroll
                for (int n1 = 0; n1 < Q; ++n1) {
                    const int n = n2 + BIG_M * n1;
                    const double re = (r1 && n < a.f1.mx) ? a.f1.p[(size_t)row * a.f1.mx + n] : 0.0;
                    const double im = (r2 && n < a.f2.mx) ? a.f2.p[(size_t)row * a.f2.mx + n] : 0.0;
                    const double2 x = cmk(re, im);
                    acc = cadd(acc, cls ? cmul(x, a.wl2x[((unsigned)n * cls) & mask]) : x);
                }
                v[e] = acc;
            }
            fft_regs<BIG_M, false>(v, ctx, xb, tw, j);
#pragma unroll
            for (int e = 0; e < 16; ++e) a.S[s_index(a.g, px, k1 + Q * (j + e * T), row)] = v[e];
        }
    }
}

// ---------------------------------------------------------------- pass B / G (columns)
// work item = one 2-column tile; group p owns the y-parity p of both the forward classes and the
// inverse half transforms.  MODE 0: convolution (S -> S2); MODE 1: Ghat = Re(FFT_y(Re S)).
template <int Q, int MODE, class Ctx>
ILM_HD void passB_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T, EQ = 16 / Q;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twy);
    const int j = ctx.tid, py = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF;
    const int Lb = a.g.Ly;                                        // = Q * BIG_M
    const unsigned mask = 2u * (unsigned)Lb - 1u;
    double2* scr = a.scratch + ((size_t)block * 2 + ctx.grp) * (size_t)Lb;       // [n1][kappa]
    const int ncols = 2 * a.g.Lx;
    const int nwork = ncols / 2;
    const bool ranged = a.whi > 0;
    const int wbeg = ranged ? a.wlo : 0, wend = ranged ? (a.whi < nwork ? a.whi : nwork) : nwork;
    if (MODE == 0 && !py) ctx.arrive(BAR_FREE);
    for (int w = wbeg + block; w < wend; w += nblocks) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int c = 2 * w + sub;
            const int px = c / a.g.Lx, m = c % a.g.Lx;
            const size_t gbase = ((size_t)ghat_col(a.g, px, m) * 2 + py) * (size_t)Lb;
            double2 v[16];
#pragma unroll 1
            for (int k1 = 0; k1 < Q; ++k1) {
                const unsigned cls = (unsigned)(py + 2 * k1);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int n2 = j + e * T;
                    double2 acc = cmk(0.0, 0.0);
#pragma unroll
                    for (int n1 = 0; n1 < Q; ++n1) {
                        const int n = n2 + BIG_M * n1;
                        if (n >= a.rlo && n < a.rhi) {
                            double2 x = a.S[s_index(a.g, px, m, n)];
                            if (MODE == 1) x.y = 0.0;
                            acc = cadd(acc, cls ? cmul(x, a.wl2y[((unsigned)n * cls) & mask]) : x);
                        }
                    }
                    v[e] = acc;
                }
                fft_regs<BIG_M, false>(v, ctx, xb, tw, j);
                if constexpr (MODE == 1) {
                    if (ghat_is_rep(a.g, px, m)) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) a.GhatOut[gbase + k1 + Q * (j + e * T)] = v[e].x * a.gscale;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const double gh = a.Ghat[gbase + k1 + Q * (j + e * T)];
                        v[e] = cmk(v[e].x * gh, v[e].y * gh);
                    }
                    // thread-local radix-Q step towards the inverse: registers e' + EQ k1' hold Z[k2 + 2M k1']
#pragma unroll
                    for (int ep = 0; ep < EQ; ++ep) {
                        const int t = j + ep * T;                                  // < M / Q
                        const unsigned k2 = cls + 2u * Q * (unsigned)t;            // < 2M
                        const int kappa = k1 + Q * t;
#pragma unroll
                        for (int n1 = 0; n1 < Q; ++n1) {
                            double2 s = v[ep];
#pragma unroll
                            for (int q = 1; q < Q; ++q) s = cadd(s, rotq<Q>(v[ep + EQ * q], n1 * q));
                            scr[(size_t)n1 * BIG_M + kappa] = n1 ? cmulc(s, a.wl2y[((unsigned)n1 * k2) & mask]) : s;
                        }
                    }
                }
            }
            if constexpr (MODE == 0) {
                ctx.sync();                                   // this group's scratch line is complete
#pragma unroll 1
                for (int n1 = 0; n1 < Q; ++n1) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = scr[(size_t)n1 * BIG_M + j + e * T];
                    fft_regs<BIG_M, true>(v, ctx, xb, tw, j);
                    if (py) {
                        ctx.wait(BAR_FREE);
#pragma unroll
                        for (int e = 0; e < 16; ++e) comb[j + e * T] = cmulc(v[e], mod_fwd<BIG_M>(tw, j, e));
                        ctx.arrive(BAR_READY);
                    } else {
                        ctx.wait(BAR_READY);
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int n = n1 + Q * (j + e * T);
                            if (n >= a.olo && n < a.ohi) a.S2[s_index(a.g, px, m, n)] = cadd(v[e], comb[j + e * T]);
                        }
                        ctx.arrive(BAR_FREE);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- pass C (rows, inverse)
// work item = (row, n1); group p inverts the half with k2 = 2 kappa + p; group 0 combines and stores
template <int Q, class Ctx>
ILM_HD void passC_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twx);
    const int j = ctx.tid, px = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF;
    const unsigned mask = 2u * (unsigned)a.g.Lx - 1u;
    const int nwork = (a.ohi - a.olo) * Q;
    if (!px) ctx.arrive(BAR_FREE);
    // the Q residues of a row run on Q neighbouring CTAs at the same time (measured faster than back
    // to back on one CTA: the spectrum row is fetched once and hit in L2 by the other three)
    for (int w = block; w < nwork; w += nblocks) {
        const int row = a.olo + w / Q, n1 = w % Q;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
        double2 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kappa = j + e * T;
            double2 s = a.S2[s_index(a.g, px, kappa, row)];
#pragma unroll
            for (int q = 1; q < Q; ++q) s = cadd(s, rotq<Q>(a.S2[s_index(a.g, px, kappa + BIG_M * q, row)], n1 * q));
            const unsigned k2 = 2u * (unsigned)kappa + (unsigned)px;
            v[e] = n1 ? cmulc(s, a.wl2x[((unsigned)n1 * k2) & mask]) : s;
        }
        fft_regs<BIG_M, true>(v, ctx, xb, tw, j);
        if (px) {
            ctx.wait(BAR_FREE);
#pragma unroll
            for (int e = 0; e < 16; ++e) comb[j + e * T] = cmulc(v[e], mod_fwd<BIG_M>(tw, j, e));
            ctx.arrive(BAR_READY);
        } else {
            ctx.wait(BAR_READY);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = n1 + Q * (j + e * T);
                const double2 y = cadd(v[e], comb[j + e * T]);
                if (r1 && n < a.f1.mx) a.f1.p[(size_t)row * a.f1.mx + n] = y.x;
                if (r2 && n < a.f2.mx) a.f2.p[(size_t)row * a.f2.mx + n] = y.y;
            }
            ctx.arrive(BAR_FREE);
        }
    }
}

}  // namespace ilm

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
// The row passes (A, C) use exactly these two formulas.  The column pass (B) decimates in TIME on the
// way in instead and spreads a column over a thread-block cluster of Q CTAs (see passB_big_body): the
// hand-off between the forward sub-transforms and the inverse residues -- the only intermediate that
// exceeds shared memory -- is one L2-resident line of 2L complex per cluster.
//
// Pass A re-reads a row Q times (real input, L2 hits); pass C gathers each spectrum row once per cluster.
//
// Round 2: the column pass of a convolution runs as THREE independent kernels per chunk of columns (passB1_big_body,
// k_big_filter_Q*, passB2_big_body<FILTERED>: forward sub-transforms -> L2-resident hand-off block -> streaming Q x Q
// filter -> inverse half transforms) on a spectrum whose rows are stored class by class (ConvGeom::rq), see below; the
// cluster form (passB_big_body) still builds the multiplier (MODE 1) and serves ILM_BIG_CHUNK=0.
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
            // class twiddle w_2L^{n cls}, n = j + e T + M n1: one table entry per thread (w^{j cls}) times warp-uniform
            // entries (w^{M cls n1}, w^{T cls e}) instead of one scattered table load per element (64 per step)
            double2 wjm[Q];
            if (cls) {
                const double2 wj = a.wl2x[((unsigned)j * cls) & mask];
#pragma unroll
                for (int n1 = 0; n1 < Q; ++n1) wjm[n1] = n1 ? cmul(wj, a.wl2x[((unsigned)(BIG_M * n1) * cls) & mask]) : wj;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n2 = j + e * T;
                const double2 wte = (cls && e) ? a.wl2x[((unsigned)(e * T) * cls) & mask] : cmk(1.0, 0.0);
                double2 acc = cmk(0.0, 0.0);
#pragma unroll
                for (int n1 = 0; n1 < Q; ++n1) {
                    const int n = n2 + BIG_M * n1;
                    const double re = (r1 && n < a.f1.mx) ? a.f1.p[(size_t)row * a.f1.mx + n] : 0.0;
                    const double im = (r2 && n < a.f2.mx) ? a.f2.p[(size_t)row * a.f2.mx + n] : 0.0;
                    const double2 x = cmk(re, im);
                    acc = cadd(acc, cls ? cmul(x, e ? cmul(wjm[n1], wte) : wjm[n1]) : x);
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
// One thread-block CLUSTER of Q CTAs per 2-column tile (columns visited one after the other), so that
// the live data of a column (input, hand-off line, output: ~1.5 MB at L = 16384) exists 148/Q times
// instead of 148 times and stays L2-resident.  Decimation in time on the way in: CTA `rank` transforms
// the sub-sequence x[rank + Q n2] (group p with the parity modulation w_2M^{n2 p}, exactly the half
// transform of ilm_conv.cuh), which reads every spectrum entry once per parity:
//     A_{n1,p}[kappa] = FFT_M( x[n1 + Q n2] w_2M^{n2 p} )[kappa]            -> hand-off line [p][n1][kappa]
// After one cluster barrier CTA `rank` = n1' builds, for k2 = 2 kappa + p,
//     Z[k2 + 2M k1]  = sum_{n1} w_2L^{n1 k2} e^{-2 pi i n1 k1 / Q} A_{n1,p}[kappa]
//     U_{n1'}[k2]    = conj(w_2L^{n1' k2}) sum_{k1} Ghat[k2 + 2M k1] Z[k2 + 2M k1] e^{+2 pi i n1' k1 / Q}
// (a Q x Q filter per frequency, thread-local) and inverts it with the usual even/odd pair of 4096-point
// transforms: z[n1' + Q n2] = Y_0[n2] + conj(w_2M^{n2}) Y_1[n2].  A second cluster barrier frees the line.
// MODE 0: convolution (S -> S2); MODE 1: Ghat[.., k2 + 2M rank] = Re Z (multiplier construction).
// w_2L^{2 T e} with L = Q * 4096, T = 256: exp(-2 pi i e / (16 Q)), e = 0..15 (a 32nd / 64th root of unity)
template <int Q> ILM_HD double2 wstep(int e) {
    if constexpr (Q == 2) return w32(e);
    else {
        // exp(-2 pi i e / 64) = w32(e / 2) * (e odd ? exp(-2 pi i / 64) : 1)
        const double2 h = w32(e >> 1);
        return (e & 1) ? cmul(h, cmk(0.99518472667219688624, -0.09801714032956060199)) : h;
    }
}

template <int Q> ILM_HD void radixq_fwd(double2* t) {      // t[k1] <- sum_{n1} t[n1] e^{-2 pi i n1 k1 / Q}
    if constexpr (Q == 2) {
        const double2 a = t[0], b = t[1];
        t[0] = cadd(a, b); t[1] = csub(a, b);
    } else {
        const double2 s02 = cadd(t[0], t[2]), d02 = csub(t[0], t[2]);
        const double2 s13 = cadd(t[1], t[3]), d13 = cmi<false>(csub(t[1], t[3]));      // -i (t1 - t3)
        t[0] = cadd(s02, s13); t[1] = cadd(d02, d13); t[2] = csub(s02, s13); t[3] = csub(d02, d13);
    }
}

template <int Q, int MODE, class Ctx>
ILM_HD void passB_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int cluster, int nclusters) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twy);
    const int j = ctx.tid, py = ctx.grp, rank = ctx.cluster_rank();
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF;
    const int Lb = a.g.Ly;                                        // = Q * BIG_M
    const unsigned mask = 2u * (unsigned)Lb - 1u;
    double2* scr = a.scratch + (size_t)cluster * 2 * (size_t)Lb + (size_t)py * Lb;      // this parity: [n1][kappa]
    const int nwork = a.g.Lx;                                     // 2-column tiles
    const bool ranged = a.whi > 0;
    const int wbeg = ranged ? a.wlo : 0, wend = ranged ? (a.whi < nwork ? a.whi : nwork) : nwork;
    if (MODE == 0 && !py) ctx.arrive(BAR_FREE);
    for (int w = wbeg + cluster; w < wend; w += nclusters) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int c = 2 * w + sub;
            const int px = c / a.g.Lx, m = c % a.g.Lx;
            const size_t gbase = ((size_t)ghat_col(a.g, px, m) * 2 + py) * (size_t)Lb;
            double2 v[16];
            // L2 prefetch, one cluster-wide share per CTA: the multiplier slice of this column (read after the
            // forward transform) and, once per tile, the spectrum of the next tile of this cluster
            if (MODE == 0) {
                const char* gp = reinterpret_cast<const char*>(a.Ghat + gbase) + (size_t)rank * BIG_M * sizeof(double);
                for (size_t off = (size_t)j * 128; off < BIG_M * sizeof(double); off += 256 * 128) ctx.prefetch_l2(gp + off);
            }
            if (sub == 0 && w + nclusters < wend && a.rhi - a.rlo == a.g.MYp) {
                const size_t share = (size_t)2 * a.g.MYp * sizeof(double2) / Q;
                const char* sp = reinterpret_cast<const char*>(a.S + (size_t)(w + nclusters) * 2 * a.g.MYp) + (size_t)rank * share;
                for (size_t off = (size_t)(ctx.grp * 256 + j) * 128; off < share; off += 512 * 128) ctx.prefetch_l2(sp + off);
            }
            // ---- forward: the decimated sub-sequence of this CTA
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = rank + Q * (j + e * T);
                double2 x = (n >= a.rlo && n < a.rhi) ? a.S[s_index(a.g, px, m, n)] : cmk(0.0, 0.0);
                if (MODE == 1) x.y = 0.0;
                v[e] = py ? cmul(x, mod_fwd<BIG_M>(tw, j, e)) : x;
            }
            fft_regs<BIG_M, false>(v, ctx, xb, tw, j);
#pragma unroll
            for (int e = 0; e < 16; ++e) scr[(size_t)rank * BIG_M + j + e * T] = v[e];
            ctx.cluster_sync();                               // all Q sub-spectra of this column are in the line
            // ---- Q x Q filter per frequency, then the inverse half transform of residue `rank`
            // twiddles w_2L^{n1 k2}, k2 = 2 (j + e T) + p: one table entry per thread (e = 0), the step
            // w_2L^{2T} = w_{L/T}... is a compile-time root of unity, the powers n1 = 2, 3 by products
            const double2 wbase = a.wl2y[(2u * (unsigned)j + (unsigned)py) & mask];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int kappa = j + e * T;
                double2 wp[Q];                                // wp[n1] = w_2L^{n1 k2}
                if constexpr (Q == 2) {                       // one twiddle per frequency: the table load is cheaper (measured)
                    wp[1] = a.wl2y[(2u * (unsigned)kappa + (unsigned)py) & mask];
                } else {
                    wp[1] = e ? cmul(wbase, wstep<Q>(e)) : wbase;
                    wp[2] = cmul(wp[1], wp[1]); wp[3] = cmul(wp[2], wp[1]);
                }
                double2 t[Q];
                t[0] = scr[kappa];
#pragma unroll
                for (int n1 = 1; n1 < Q; ++n1) t[n1] = cmul(scr[(size_t)n1 * BIG_M + kappa], wp[n1]);
                radixq_fwd<Q>(t);                             // t[k1] = Z[k2 + 2M k1]
                if constexpr (MODE == 1) {
                    if (ghat_is_rep(a.g, px, m)) a.GhatOut[gbase + kappa + (size_t)BIG_M * rank] = t[rank].x * a.gscale;
                } else {
                    double2 s = cmk(0.0, 0.0);
#pragma unroll
                    for (int k1 = 0; k1 < Q; ++k1) {
                        const double gh = a.Ghat[gbase + kappa + (size_t)BIG_M * k1];
                        s = cadd(s, rotq<Q>(cmk(t[k1].x * gh, t[k1].y * gh), rank * k1));
                    }
                    v[e] = rank ? cmulc(s, wp[rank]) : s;
                }
            }
            if constexpr (MODE == 0) {
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
                        const int n = rank + Q * (j + e * T);
                        if (n >= a.olo && n < a.ohi) a.S2[s_index(a.g, px, m, n)] = cadd(v[e], comb[j + e * T]);
                    }
                    ctx.arrive(BAR_FREE);
                }
            }
            ctx.cluster_sync();                               // the line may be overwritten by the next column
        }
    }
}

// ---------------------------------------------------------------- pass B split in two kernels (the default)
// The cluster form above couples Q CTAs through two cluster barriers per column and exposes every latency of a column
// (decimated loads, hand-off line, multiplier) to all of them: at 16384^2 it runs at 17 % issue utilisation, 50 %
// long-scoreboard stalls.  Here the two halves are separate kernels over independent work items (column, residue):
//   B1  A_{n1,p}[kappa] = FFT_M( x[n1 + Q n2] w_2M^{n2 p} )[kappa]        -> bigA[col - c0][p][n1][kappa]
//   B2  the Q x Q filter per frequency and the inverse half transform of residue n1' (as above)  -> S2
// for a chunk [c0, c0 + nc) of columns at a time, sized so that its hand-off block (nc * 2L complex) stays L2-resident
// between the two launches.  The 2Q items of a tile run on neighbouring CTAs at the same time, so the 128-byte lines of
// the tile (16 bytes per item) come from DRAM once.
template <int Q, class Ctx>
ILM_HD void passB1_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twy);
    const int j = ctx.tid, py = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    const int Lb = a.g.Ly;
    const int nitems = a.bnc * Q;
    for (int it = block; it < nitems; it += nblocks) {
        const int c = a.bc0 + it / Q, rank = it % Q;
        const int px = c / a.g.Lx, m = c % a.g.Lx;
        const int itn = it + nblocks;                              // L2 prefetch: this item's eighth of the tile of the next item
        if (itn < nitems && a.rhi - a.rlo == a.g.MYp) {
            const int cn = a.bc0 + itn / Q;
            const size_t share = (size_t)2 * a.g.MYp * sizeof(double2) / (2 * Q);
            const char* sp = reinterpret_cast<const char*>(a.S + (size_t)(cn >> 1) * 2 * a.g.MYp) + (size_t)((cn & 1) * Q + itn % Q) * share;
            for (size_t off = (size_t)(ctx.grp * 256 + j) * 128; off < share; off += 512 * 128) ctx.prefetch_l2(sp + off);
        }
        double2 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int n = rank + Q * (j + e * T);
            const double2 x = (n >= a.rlo && n < a.rhi) ? a.S[s_index(a.g, px, m, n)] : cmk(0.0, 0.0);
            v[e] = py ? cmul(x, mod_fwd<BIG_M>(tw, j, e)) : x;
        }
        fft_regs<BIG_M, false>(v, ctx, xb, tw, j);
        double2* dst = a.bigA + ((size_t)(c - a.bc0) * 2 + py) * (size_t)Lb + (size_t)rank * BIG_M;
#pragma unroll
        for (int e = 0; e < 16; ++e) dst[j + e * T] = v[e];
    }
}

// FILTERED: the Q x Q filter was applied in place by k_big_filter (a streaming kernel at full occupancy; inside this
// kernel the 8 dependent loads per frequency are exposed 16 times per item: 72 % long-scoreboard stalls), the hand-off
// block then holds U_{n1'}[k2] at [p][n1'][kappa] and an item loads its 16 entries per thread like any half transform.
template <int Q, bool FILTERED, class Ctx>
ILM_HD void passB2_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twy);
    const int j = ctx.tid, py = ctx.grp;
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF;
    const int Lb = a.g.Ly;
    const unsigned mask = 2u * (unsigned)Lb - 1u;
    const int nitems = a.bnc * Q;
    const double2 wbase = a.wl2y[(2u * (unsigned)j + (unsigned)py) & mask];
    if (!py) ctx.arrive(BAR_FREE);
    for (int it = block; it < nitems; it += nblocks) {
        const int c = a.bc0 + it / Q, rank = it % Q;
        const int px = c / a.g.Lx, m = c % a.g.Lx;
        const size_t gbase = ((size_t)ghat_col(a.g, px, m) * 2 + py) * (size_t)Lb;
        const int itn = it + nblocks;                              // L2 prefetch of the next item's multiplier share
        if (!FILTERED && itn < nitems) {
            const int cn = a.bc0 + itn / Q;
            const char* gp = reinterpret_cast<const char*>(a.Ghat + ((size_t)ghat_col(a.g, cn / a.g.Lx, cn % a.g.Lx) * 2 + py) * (size_t)Lb) +
                             (size_t)(itn % Q) * BIG_M * sizeof(double);
            for (size_t off = (size_t)j * 128; off < BIG_M * sizeof(double); off += 256 * 128) ctx.prefetch_l2(gp + off);
        }
        const double2* scr = a.bigA + ((size_t)(c - a.bc0) * 2 + py) * (size_t)Lb;
        double2 v[16];
        if constexpr (FILTERED) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = scr[(size_t)rank * BIG_M + j + e * T];
        } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kappa = j + e * T;
            double2 wp[Q];                                // wp[n1] = w_2L^{n1 k2}, k2 = 2 kappa + py
            if constexpr (Q == 2) {
                wp[1] = a.wl2y[(2u * (unsigned)kappa + (unsigned)py) & mask];
            } else {
                wp[1] = e ? cmul(wbase, wstep<Q>(e)) : wbase;
                wp[2] = cmul(wp[1], wp[1]); wp[3] = cmul(wp[2], wp[1]);
            }
            double2 t[Q];
            t[0] = scr[kappa];
#pragma unroll
            for (int n1 = 1; n1 < Q; ++n1) t[n1] = cmul(scr[(size_t)n1 * BIG_M + kappa], wp[n1]);
            radixq_fwd<Q>(t);                             // t[k1] = Z[k2 + 2M k1]
            double2 s = cmk(0.0, 0.0);
#pragma unroll
            for (int k1 = 0; k1 < Q; ++k1) {
                const double gh = a.Ghat[gbase + kappa + (size_t)BIG_M * k1];
                s = cadd(s, rotq<Q>(cmk(t[k1].x * gh, t[k1].y * gh), rank * k1));
            }
            v[e] = rank ? cmulc(s, wp[rank]) : s;
        }
        }
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
                const int n = rank + Q * (j + e * T);
                if (n >= a.olo && n < a.ohi) a.S2[s_index(a.g, px, m, n)] = cadd(v[e], comb[j + e * T]);
            }
            ctx.arrive(BAR_FREE);
        }
    }
}

// the Q x Q filter of one frequency, in place: t[n1] = A_{n1,p}[kappa]  ->  t[n1'] = U_{n1'}[k2], k2 = 2 kappa + p
//     Z[k2 + 2M k1] = sum_{n1} w_2L^{n1 k2} e^{-2 pi i n1 k1 / Q} A_{n1},   U_{n1'} = conj(w_2L^{n1' k2}) sum_{k1} Ghat[k2 + 2M k1] Z[..] e^{+2 pi i n1' k1 / Q}
// (same operations, same order as the in-kernel filter above: bit-identical results)
template <int Q> ILM_HD void big_filter_point(double2* t, const double* gh, const double2* wp) {
#pragma unroll
    for (int n1 = 1; n1 < Q; ++n1) t[n1] = cmul(t[n1], wp[n1]);
    radixq_fwd<Q>(t);
    double2 z[Q];
#pragma unroll
    for (int k1 = 0; k1 < Q; ++k1) z[k1] = cmk(t[k1].x * gh[k1], t[k1].y * gh[k1]);
#pragma unroll
    for (int r = 0; r < Q; ++r) {
        double2 s = cmk(0.0, 0.0);
#pragma unroll
        for (int k1 = 0; k1 < Q; ++k1) s = cadd(s, rotq<Q>(z[k1], r * k1));
        t[r] = r ? cmulc(s, wp[r]) : s;
    }
}

// ---------------------------------------------------------------- pass C (rows, inverse), one CTA per (row, residue)
// Used for Q = 2 (measured faster there than the cluster version below: 2.17 vs 2.80 ms at 8192^2).
// work item = (row, n1); group p inverts the half with k2 = 2 kappa + p; group 0 combines and stores
template <int Q, class Ctx>
ILM_HD void passC_big_percta_body(Ctx& ctx, const ConvArgs& a, double2* smem, int block, int nblocks) {
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

// ---------------------------------------------------------------- pass C (rows, inverse), one cluster per row (Q = 4)
// One cluster of Q CTAs per row.  CTA `rank` gathers block k1 = rank of the row's spectrum (the 32-byte
// pieces of the 2x2 tiles, one tile column apart) ONCE into the cluster's L2-resident line; after a
// cluster barrier CTA `rank` = n1 reads the Q blocks contiguously, forms
//     U_{n1}[k2] = conj(w_2L^{n1 k2}) sum_{k1} Z[k2 + 2M k1] e^{+2 pi i n1 k1 / Q},   k2 = 2 kappa + p,
// and inverts its residue: z[n1 + Q n2] = Y_0[n2] + conj(w_2M^{n2}) Y_1[n2] (group p = parity, combined
// through shared memory as in ilm_conv.cuh).  Without the line every residue re-gathered the whole row.
template <int Q, class Ctx>
ILM_HD void passC_big_body(Ctx& ctx, const ConvArgs& a, double2* smem, int cluster, int nclusters) {
    using C = FftCfg<BIG_M>;
    constexpr int T = C::T;
    double2* tw = smem + C::TW_BASE;
    load_twiddles<BIG_M>(ctx, tw, a.twx);
    const int j = ctx.tid, px = ctx.grp, rank = ctx.cluster_rank();
    double2* xb = smem + ctx.grp * C::GROUP_XBUF;
    double2* comb = smem + 2 * C::GROUP_XBUF;
    const int Lb = a.g.Lx;
    const unsigned mask = 2u * (unsigned)Lb - 1u;
    double2* line = a.scratch + (size_t)cluster * 2 * (size_t)Lb + (size_t)px * Lb;      // this parity: [k1][kappa]
    const int nrows = a.ohi - a.olo;
    if (!px) ctx.arrive(BAR_FREE);
    const double2 wbase = a.wl2x[(2u * (unsigned)j + (unsigned)px) & mask];
    for (int w = cluster; w < nrows; w += nclusters) {
        const int row = a.olo + w;
        const bool r1 = a.f1.p && row < a.f1.my, r2 = a.f2.p && row < a.f2.my;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kappa = j + e * T;
            line[(size_t)rank * BIG_M + kappa] = a.S2[s_index(a.g, px, kappa + BIG_M * rank, row)];
        }
        ctx.cluster_sync();                                   // the whole row is in the line
        double2 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int kappa = j + e * T;
            double2 s = line[kappa];
#pragma unroll
            for (int q = 1; q < Q; ++q) s = cadd(s, rotq<Q>(line[(size_t)q * BIG_M + kappa], rank * q));
            if (rank) {
                double2 wp = e ? cmul(wbase, wstep<Q>(e)) : wbase;            // w_2L^{k2}
                if (Q == 4 && rank >= 2) { const double2 w2 = cmul(wp, wp); wp = rank == 3 ? cmul(w2, wp) : w2; }
                s = cmulc(s, wp);
            }
            v[e] = s;
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
                const int n = rank + Q * (j + e * T);
                const double2 y = cadd(v[e], comb[j + e * T]);
                if (r1 && n < a.f1.mx) a.f1.p[(size_t)row * a.f1.mx + n] = y.x;
                if (r2 && n < a.f2.mx) a.f2.p[(size_t)row * a.f2.mx + n] = y.y;
            }
            ctx.arrive(BAR_FREE);
        }
        ctx.cluster_sync();                                   // the line may be overwritten by the next row
    }
}

}  // namespace ilm

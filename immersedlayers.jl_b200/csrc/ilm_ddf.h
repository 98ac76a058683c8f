// ilm_ddf.h -- discrete delta functions and the per-point window table entry,
// shared by host and device so that index/weight tables are bit-reproducible.
//
// Replaces CartesianGrids' DDF / Regularize functor that ImmersedLayers
// instantiates in _get_regularization (src/cache.jl:305-314) and turns into
// sparse R/E matrices in _regularization_matrix / _interpolation_matrix
// (src/cache.jl:328-347).  Formulas: SURVEY.md A.2.
//
// The operation ORDER in every expression below is part of the contract with
// the test oracle (oracle/ilm_oracle.py): both evaluate the same sequence of
// IEEE-754 +,-,*,/,sqrt, with no fused multiply-add.  This translation unit must
// therefore be compiled with --fmad=false (device) / -ffp-contract=off (host).
// asin follows the published fdlibm e_asin.c algorithm (the one Julia's
// Base.asin implements) so that host, device and oracle agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ILM_DDF_HD __host__ __device__ inline
#else
#define ILM_DDF_HD inline
#endif

namespace ilm {

enum DdfKind { DDF_YANG3 = 0, DDF_M3 = 1, DDF_ROMA = 2, DDF_M4PRIME = 3, DDF_WITCHHAT = 4, DDF_COUNT = 5 };

ILM_DDF_HD int ddf_width(int kind) {
    return (kind == DDF_YANG3 || kind == DDF_M4PRIME) ? 4 : (kind == DDF_WITCHHAT ? 2 : 3);
}
ILM_DDF_HD double ddf_radius(int kind) {
    return (kind == DDF_YANG3 || kind == DDF_M4PRIME) ? 2.0 : (kind == DDF_WITCHHAT ? 1.0 : 1.5);
}

ILM_DDF_HD double ilm_hi_word_trunc(double s) {   // SET_LOW_WORD(w, 0)
    uint64_t u;
    memcpy(&u, &s, 8);
    u &= 0xFFFFFFFF00000000ull;
    double w;
    memcpy(&w, &u, 8);
    return w;
}
ILM_DDF_HD int32_t ilm_hi_word(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int32_t)(u >> 32);
}

ILM_DDF_HD double asin_fdlibm(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
                 pio4_hi = 7.85398163397448278999e-01;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double ax = fabs(x);
    if (ax < 0.5) {
        if (ax < 7.450580596923828125e-09) return x;     // 2^-27
        const double t = x * x;
        const double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
        const double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
        const double w = p / q;
        return x + x * w;
    }
    if (ax >= 1.0) {
        const double r = ax * pio2_hi + ax * pio2_lo;
        return x > 0 ? r : -r;
    }
    double w = 1.0 - ax;
    double t = w * 0.5;
    double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
    double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
    const double s = sqrt(t);
    if ((ilm_hi_word(ax) & 0x7fffffff) >= 0x3FEF3333) {
        w = p / q;
        t = pio2_hi - (2.0 * (s + s * w) - pio2_lo);
    } else {
        w = ilm_hi_word_trunc(s);
        const double c = (t - w * w) / (s + w);
        const double r = p / q;
        p = 2.0 * s * r - (pio2_lo - 2.0 * c);
        q = pio4_hi - 2.0 * w;
        t = pio4_hi - (p - q);
    }
    return x > 0 ? t : -t;
}

ILM_DDF_HD double ddf_eval(int kind, double rs) {
    const double r = fabs(rs);
    switch (kind) {
    case DDF_YANG3: {
        const double C1 = 0.40454998233983938, C2 = 1.0954500176601606, S12 = 0.14433756729740644,
                     S2 = 0.86602540378443865, S36 = 0.048112522432468814, C1312 = 1.0833333333333333,
                     C148 = 0.020833333333333332;
        if (r <= 1.0) {
            const double rr = r * r;
            const double s = sqrt(((-12.0 * rr) + 12.0 * r) + 1.0);
            const double t3 = ((1.0 - 2.0 * r) * 0.0625) * s;
            const double t4 = S12 * asin_fdlibm(S2 * (2.0 * r - 1.0));
            return (((C1 + r * 0.25) - rr * 0.25) + t3) - t4;
        }
        if (r < 2.0) {
            const double rr = r * r;
            const double s = sqrt(((-12.0 * rr) + 36.0 * r) - 23.0);
            const double t3 = ((2.0 * r - 3.0) * C148) * s;
            const double t4 = S36 * asin_fdlibm(S2 * (2.0 * r - 3.0));
            return (((C2 - C1312 * r) + rr * 0.25) + t3) + t4;
        }
        return 0.0;
    }
    case DDF_M3:
        if (r <= 0.5) return 0.75 - r * r;
        if (r < 1.5) { const double d = 1.5 - r; return 0.5 * (d * d); }
        return 0.0;
    case DDF_ROMA:
        if (r <= 0.5) return (1.0 + sqrt(1.0 - 3.0 * (r * r))) / 3.0;
        if (r < 1.5) { const double d = 1.0 - r; return ((5.0 - 3.0 * r) - sqrt(1.0 - 3.0 * (d * d))) / 6.0; }
        return 0.0;
    case DDF_M4PRIME:
        if (r <= 1.0) { const double rr = r * r; return (1.0 - 2.5 * rr) + 1.5 * (rr * r); }
        if (r < 2.0) { const double d = 2.0 - r; return (0.5 * (d * d)) * (1.0 - r); }
        return 0.0;
    default:   // DDF_WITCHHAT
        return r < 1.0 ? 1.0 - r : 0.0;
    }
}

// Window table of one surface point for one target layout.
//   shift (sx, sy): index-space offset of the layout
//   field extents (mx, my); wgt = ds/dx^2 (GridScaling) or 1 (IndexScaling)
// Outputs: i0, j0 = 0-based array index of the first window entry;
//   wR[b*W+a], wE[b*W+a] for a (x), b (y) in [0,W), zero where the entry falls
//   outside the field.
ILM_DDF_HD void ddf_point_table(int kind, double xs, double ys, double wgt, bool symmetric, double sx, double sy,
                                int mx, int my, int* i0, int* j0, double* wR, double* wE) {
    const int W = ddf_width(kind);
    const double rho = ddf_radius(kind);
    const int ib = (int)floor((xs + sx) - rho) + 1;    // 1-based
    const int jb = (int)floor((ys + sy) - rho) + 1;
    double px[4], py[4];
    for (int a = 0; a < W; ++a) {
        px[a] = ddf_eval(kind, ((double)(ib + a) - sx) - xs);
        py[a] = ddf_eval(kind, ((double)(jb + a) - sy) - ys);
    }
    for (int b = 0; b < W; ++b)
        for (int a = 0; a < W; ++a) {
            const int i1 = ib + a, j1 = jb + b;
            const bool valid = i1 >= 1 && i1 <= mx && j1 >= 1 && j1 <= my;
            const double e = px[a] * py[b];
            const double r = e * wgt;
            wE[b * W + a] = valid ? (symmetric ? r : e) : 0.0;
            wR[b * W + a] = valid ? r : 0.0;
        }
    *i0 = ib - 1;
    *j0 = jb - 1;
}

}  // namespace ilm

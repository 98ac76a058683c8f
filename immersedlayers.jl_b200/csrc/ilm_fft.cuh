// ilm_fft.cuh -- register/shared-memory fp64 FFT building blocks for the
// lattice-Green's-function convolution (row a7/a8 of SURVEY.md section 8).
//
// Replaces the FFTW rfft/irfft calls inside CartesianGrids' CircularConvolution
// reached from ImmersedLayers' inverse_laplacian! (src/grid_operators.jl:153-179).
//
// Design:
//   * one "group" = 256 threads (8 warps) owns F = 4096/L complex FFTs of
//     length L at a time, 16 elements per thread in registers;
//   * Stockham autosort passes of radix 16 (plus one radix-2/4/8 tail pass),
//     element e of thread j is always x[j + e*T], T = L/16, both before the
//     first pass and after the last one, so global loads/stores are coalesced
//     straight from/to registers;
//   * between passes the group exchanges through a padded shared-memory buffer
//     (index i -> i + i/16, conflict-free for 128-bit accesses);
//   * twiddles come from a shared-memory table laid out [t][k] (conflict-free).
// Every function is __host__ __device__ and templated on a context type so
// that the exact kernel bodies can be executed on the CPU by the emulation
// harness in csrc/test_fft_host.cu (no GPU in the build container).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ILM_HD __host__ __device__ __forceinline__
#else
#define ILM_HD inline
#endif

namespace ilm {

ILM_HD double2 cmk(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
ILM_HD double2 cadd(double2 a, double2 b) { return cmk(a.x + b.x, a.y + b.y); }
ILM_HD double2 csub(double2 a, double2 b) { return cmk(a.x - b.x, a.y - b.y); }
ILM_HD double2 cmul(double2 a, double2 b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
ILM_HD double2 cmulc(double2 a, double2 b) { return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// multiply by the twiddle tw (forward) or its conjugate (inverse)
template <bool INV> ILM_HD double2 ctw(double2 a, double2 tw) { return INV ? cmulc(a, tw) : cmul(a, tw); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> ILM_HD double2 cmi(double2 a) { return INV ? cmk(-a.y, a.x) : cmk(a.y, -a.x); }

#define ILM_SQRT1_2 0.70710678118654752440
#define ILM_COS_PI_8 0.92387953251128675613
#define ILM_SIN_PI_8 0.38268343236508977173

// ---- radix kernels on strided register arrays, natural-order outputs -------
template <int S> ILM_HD void fft2s(double2* v) {
    double2 a = v[0], b = v[S];
    v[0] = cadd(a, b);
    v[S] = csub(a, b);
}

template <bool INV, int S> ILM_HD void fft4s(double2* v) {
    double2 s02 = cadd(v[0], v[2 * S]), d02 = csub(v[0], v[2 * S]);
    double2 s13 = cadd(v[S], v[3 * S]), d13 = cmi<INV>(csub(v[S], v[3 * S]));
    v[0] = cadd(s02, s13);
    v[S] = cadd(d02, d13);
    v[2 * S] = csub(s02, s13);
    v[3 * S] = csub(d02, d13);
}

// (x + iy) * (h - ih) forward, (h + ih) inverse, h = sqrt(1/2)
template <bool INV> ILM_HD double2 cw8_1(double2 a) {
    return INV ? cmk((a.x - a.y) * ILM_SQRT1_2, (a.x + a.y) * ILM_SQRT1_2)
               : cmk((a.x + a.y) * ILM_SQRT1_2, (a.y - a.x) * ILM_SQRT1_2);
}
// w8^3 = (-h, -h) forward, (-h, +h) inverse
template <bool INV> ILM_HD double2 cw8_3(double2 a) {
    return INV ? cmk(-(a.x + a.y) * ILM_SQRT1_2, (a.x - a.y) * ILM_SQRT1_2)
               : cmk((a.y - a.x) * ILM_SQRT1_2, -(a.x + a.y) * ILM_SQRT1_2);
}

template <bool INV, int S> ILM_HD void fft8s(double2* v) {
    // n = c + 2d, k = k1 + 4 k2
    double2 y0[4] = {v[0], v[2 * S], v[4 * S], v[6 * S]};
    double2 y1[4] = {v[S], v[3 * S], v[5 * S], v[7 * S]};
    fft4s<INV, 1>(y0);
    fft4s<INV, 1>(y1);
    y1[1] = cw8_1<INV>(y1[1]);
    y1[2] = cmi<INV>(y1[2]);
    y1[3] = cw8_3<INV>(y1[3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        v[k1 * S] = cadd(y0[k1], y1[k1]);
        v[(k1 + 4) * S] = csub(y0[k1], y1[k1]);
    }
}

template <bool INV> ILM_HD void fft16(double2* v) {
    // n = c + 4d, k = k1 + 4 k2 ; y[c][k1] = sum_d v[c+4d] w4^(d k1)
    double2 y[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        y[c][0] = v[c]; y[c][1] = v[c + 4]; y[c][2] = v[c + 8]; y[c][3] = v[c + 12];
        fft4s<INV, 1>(y[c]);
    }
    // twiddle y[c][k1] *= w16^(c k1)
    const double2 w1 = cmk(ILM_COS_PI_8, -ILM_SIN_PI_8);
    const double2 w3 = cmk(ILM_SIN_PI_8, -ILM_COS_PI_8);
    y[1][1] = ctw<INV>(y[1][1], w1);
    y[1][2] = cw8_1<INV>(y[1][2]);
    y[1][3] = ctw<INV>(y[1][3], w3);
    y[2][1] = cw8_1<INV>(y[2][1]);
    y[2][2] = cmi<INV>(y[2][2]);
    y[2][3] = cw8_3<INV>(y[2][3]);
    y[3][1] = ctw<INV>(y[3][1], w3);
    y[3][2] = cw8_3<INV>(y[3][2]);
    y[3][3] = ctw<INV>(y[3][3], cmk(-ILM_COS_PI_8, ILM_SIN_PI_8));   // w16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        double2 z[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
        fft4s<INV, 1>(z);
        v[k1] = z[0]; v[k1 + 4] = z[1]; v[k1 + 8] = z[2]; v[k1 + 12] = z[3];
    }
}

// ---- compile-time plan for length L ---------------------------------------
constexpr int ilm_log2(int n) { return n <= 1 ? 0 : 1 + ilm_log2(n / 2); }

template <int L> struct FftCfg {
    static_assert(L >= 16 && L <= 4096 && (L & (L - 1)) == 0, "L must be a power of two in [16,4096]");
    static constexpr int T = L / 16;            // threads per FFT
    static constexpr int F = 256 / T;           // FFTs per 256-thread group
    static constexpr int P = (L == 16) ? 1 : (L <= 256 ? 2 : 3);
    static constexpr int RL = (P == 1) ? 16 : (P == 2 ? L / 16 : L / 256);   // tail radix
    static constexpr int NS_LAST = L / RL;
    // twiddle tables hold only the power-of-two exponents w^(k 2^i); the other
    // powers are formed by products in registers (4 loads instead of 15)
    static constexpr int TW1_OFF = 0;
    static constexpr int TW1_N = (P == 3) ? 4 * 16 : 0;          // middle pass: radix 16, Ns = 16
    static constexpr int TWL_OFF = TW1_OFF + TW1_N;
    static constexpr int TWL_N = (P >= 2) ? ilm_log2(RL) * NS_LAST : 0;
    static constexpr int MOD_OFF = TWL_OFF + TWL_N;     // w_{2L}^j, j < T
    static constexpr int MOD_N = T;
    static constexpr int TW_TOTAL = MOD_OFF + MOD_N;
    static constexpr int XBUF = L + L / 16;             // padded exchange entries per FFT
    static constexpr int GROUP_XBUF = F * XBUF;         // = 4352 for every L
    static constexpr int COMB = F * L;                  // combine buffer (unpadded), = 4096
    static constexpr int TW_BASE = 2 * GROUP_XBUF + COMB;
    static constexpr int MBAR_OFF = TW_BASE + TW_TOTAL;      // two 8-byte mbarriers (one per group)
    static constexpr size_t SMEM_BYTES = (size_t)(MBAR_OFF + 2) * sizeof(double2);
    static constexpr bool USE_TMA = L >= 512;               // bulk-tensor loads of spectrum rows (pass C)
    static constexpr int TMA_BOX = 256;                     // tile columns (a) per bulk-tensor box
};

ILM_HD int xpad(int i) { return i + (i >> 4); }

// w_32^e, e = 0..15
ILM_HD double2 w32(int e) {
    static constexpr double c[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                          0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                          0.19509032201612826785, 0.0, -0.19509032201612826785, -0.38268343236508977173,
                          -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                          -0.92387953251128675613, -0.98078528040323044913};
    static constexpr double s[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                          0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613,
                          0.98078528040323044913, 1.0, 0.98078528040323044913, 0.92387953251128675613,
                          0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                          0.38268343236508977173, 0.19509032201612826785};
    return cmk(c[e], -s[e]);
}

// w_16^q, q = 0..15
ILM_HD double2 w16(int q) {
    const double2 h = w32((2 * q) & 15);          // w_32^(2q mod 16); w_32^16 = -1
    return (q & 8) ? cmk(-h.x, -h.y) : h;
}

template <int R, bool INV, int S> ILM_HD void fft_tail(double2* v) {
    if constexpr (R == 16) fft16<INV>(v);
    else if constexpr (R == 8) fft8s<INV, S>(v);
    else if constexpr (R == 4) fft4s<INV, S>(v);
    else fft2s<S>(v);
}

// v[t*S] *= w^(k t), t = 1..R-1, from the table of power-of-two exponents
// tw[i*ns + k] = w^(k 2^i); conjugated for the inverse transform.
template <int R, bool INV, int S> ILM_HD void twiddle_powers(double2* v, const double2* tw, int ns, int k) {
    double2 b[R];
    b[1] = tw[k];
    if constexpr (R >= 4) { b[2] = tw[ns + k]; b[3] = cmul(b[1], b[2]); }
    if constexpr (R >= 8) {
        b[4] = tw[2 * ns + k];
#pragma unroll
        for (int t = 1; t < 4; ++t) b[4 + t] = cmul(b[4], b[t]);
    }
    if constexpr (R >= 16) {
        b[8] = tw[3 * ns + k];
#pragma unroll
        for (int t = 1; t < 8; ++t) b[8 + t] = cmul(b[8], b[t]);
    }
#pragma unroll
    for (int t = 1; t < R; ++t) v[t * S] = ctw<INV>(v[t * S], b[t]);
}

// One complex FFT of length L on the 16 registers of each of its T threads.
//   v[e] holds x[j + e*T] on entry and X[j + e*T] on exit (unnormalised).
//   xb: this FFT's padded exchange buffer (XBUF entries), tw: twiddle table.
//   ctx.sync() is a barrier over the 256-thread group.
// fft_head runs every pass but the last and ends right after the last read of the
// exchange buffer (callers may then hand the buffer to an asynchronous copy);
// fft_last_pass is the last butterfly pass on registers only.
template <int L, bool INV, class Ctx>
ILM_HD void fft_head(double2* v, Ctx& ctx, double2* xb, const double2* tw, int j) {
    using C = FftCfg<L>;
    constexpr int T = C::T;
    if (C::P == 1) return;
    fft16<INV>(v);
    ctx.sync();                                   // previous readers of xb are done
#pragma unroll
    for (int t = 0; t < 16; ++t) xb[xpad(j * 16 + t)] = v[t];
    ctx.sync();
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = xb[xpad(j + e * T)];
    if (C::P == 3) {
        const int k = j & 15;
        twiddle_powers<16, INV, 1>(v, tw + C::TW1_OFF, 16, k);
        fft16<INV>(v);
        ctx.sync();
#pragma unroll
        for (int t = 0; t < 16; ++t) xb[xpad((j - k) * 16 + k + t * 16)] = v[t];
        ctx.sync();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = xb[xpad(j + e * T)];
    }
}

template <int L, bool INV> ILM_HD void fft_last_pass(double2* v, const double2* tw, int j) {
    using C = FftCfg<L>;
    if (C::P == 1) { fft16<INV>(v); return; }
    // tail pass: radix RL, Ns = L/RL, butterflies q = 0..S-1 on registers q + t*S
    constexpr int T = C::T, RL = C::RL, S = 16 / RL, NS = C::NS_LAST;
#pragma unroll
    for (int q = 0; q < S; ++q) {
        twiddle_powers<RL, INV, S>(v + q, tw + C::TWL_OFF, NS, j + q * T);
        fft_tail<RL, INV, S>(v + q);
    }
}

template <int L, bool INV, class Ctx>
ILM_HD void fft_regs(double2* v, Ctx& ctx, double2* xb, const double2* tw, int j) {
    fft_head<L, INV>(v, ctx, xb, tw, j);
    fft_last_pass<L, INV>(v, tw, j);
}

// host-side generation of the twiddle table for length L (long double -> double)
template <int L, class TrigFn>
inline void fft_fill_twiddles(double2* tw, TrigFn expm2pii /* (num, den) -> exp(-2 pi i num/den) */) {
    using C = FftCfg<L>;
    if (C::P == 3)
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 16; ++k) tw[C::TW1_OFF + i * 16 + k] = expm2pii((long long)k << i, 256);
    if (C::P >= 2)
        for (int i = 0; i < ilm_log2(C::RL); ++i)
            for (int k = 0; k < C::NS_LAST; ++k)
                tw[C::TWL_OFF + i * C::NS_LAST + k] = expm2pii((long long)k << i, (long long)L);
    for (int j = 0; j < C::T; ++j) tw[C::MOD_OFF + j] = expm2pii(j, 2LL * L);
}

}  // namespace ilm

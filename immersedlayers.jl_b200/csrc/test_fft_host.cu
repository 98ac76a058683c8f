// CPU emulation harness for the FFT-convolution kernel bodies (no GPU in the
// build container): runs the *same* __host__ __device__ bodies that the CUDA
// kernels call, with 512 std::threads per emulated CTA and condition-variable
// barriers, and checks them against a long-double DFT / direct convolution.
//   nvcc -std=c++17 -O1 -o test_fft_host test_fft_host.cu && ./test_fft_host
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "ilm_conv_big.cuh"

using namespace ilm;

struct Barrier {
    std::mutex m; std::condition_variable cv; int n, count = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
    void arrive() {          // bar.arrive: count in, do not wait
        std::unique_lock<std::mutex> lk(m);
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
    }
};

struct HostCtx {
    int tid, grp; Barrier* gb; Barrier* cb; Barrier* named;   // named[id] for id = BAR_READY, BAR_FREE
    int crank = 0; Barrier* clb = nullptr;                    // thread-block cluster: rank and cluster barrier
    int cluster_rank() const { return crank; }
    void cluster_sync() { if (clb) clb->wait(); }
    void sync() { gb->wait(); }
    void sync_cta() { cb->wait(); }
    bool any(bool) { return true; }          // a warp never skips in the emulation
    void arrive(int id) { named[id].arrive(); }
    void wait(int id) { named[id].wait(); }
    void delay(int) {}
    // bulk-tensor copies emulated synchronously by thread 0; the wait is a group barrier
    void tma_init(double2*) { gb->wait(); }
    void tma_load_rows(const ConvArgs& a, int px, int row0, int nrows, int L, int slot_stride, double2* dst, double2*) {
        if (tid != 0) return;
        for (int r = 0; r < nrows; ++r)
            for (int m = 0; m < L; ++m) {
                const int row = row0 + r;
                dst[(size_t)r * slot_stride + m] = row < a.g.MYp ? a.S2[s_index(a.g, px, m, row)] : cmk(0.0, 0.0);
            }
    }
    void tma_wait(double2*, int) { gb->wait(); }
    void prefetch_l2(const void*) {}
};

static double2 expm2pii(long long num, long long den) {
    num %= den;
    long double a = -2.0L * acosl(-1.0L) * (long double)num / (long double)den;
    return cmk((double)cosl(a), (double)sinl(a));
}

template <class Body> static void run_cta(size_t smem_bytes, Body body) {
    std::vector<double2> smem(smem_bytes / sizeof(double2) + 1);
    Barrier g0(256), g1(256), cb(512);
    Barrier named[5] = {Barrier(512), Barrier(512), Barrier(512), Barrier(512), Barrier(512)};
    std::vector<std::thread> th;
    for (int t = 0; t < 512; ++t)
        th.emplace_back([&, t] {
            HostCtx c{t & 255, t >> 8, (t >> 8) ? &g1 : &g0, &cb, named};
            body(c, smem.data());
        });
    for (auto& x : th) x.join();
}

// a thread-block cluster of Q emulated CTAs (Q x 512 threads, one shared-memory image each)
template <class Body> static void run_cluster(int Q, size_t smem_bytes, Body body) {
    struct Cta {
        std::vector<double2> smem; Barrier g0{256}, g1{256}, cb{512};
        Barrier named[5] = {Barrier(512), Barrier(512), Barrier(512), Barrier(512), Barrier(512)};
    };
    std::vector<std::unique_ptr<Cta>> ctas;
    for (int q = 0; q < Q; ++q) { ctas.emplace_back(new Cta); ctas.back()->smem.resize(smem_bytes / sizeof(double2) + 1); }
    Barrier clb(Q * 512);
    std::vector<std::thread> th;
    for (int q = 0; q < Q; ++q)
        for (int t = 0; t < 512; ++t)
            th.emplace_back([&, q, t] {
                Cta& c = *ctas[q];
                HostCtx h{t & 255, t >> 8, (t >> 8) ? &c.g1 : &c.g0, &c.cb, c.named, q, &clb};
                body(h, c.smem.data());
            });
    for (auto& x : th) x.join();
}

// ---------------------------------------------------------------- single FFT
template <int L, bool INV> static double test_fft() {
    using C = FftCfg<L>;
    std::vector<double2> tw(C::TW_TOTAL + 1);
    fft_fill_twiddles<L>(tw.data(), expm2pii);
    std::mt19937_64 rng(L + INV);
    std::normal_distribution<double> nd;
    const int nf = 2 * C::F;
    std::vector<double2> in((size_t)nf * L), out((size_t)nf * L);
    for (auto& z : in) z = cmk(nd(rng), nd(rng));
    run_cta(C::SMEM_BYTES, [&](HostCtx& c, double2* smem) {
        double2* tws = smem + C::TW_BASE;
        load_twiddles<L>(c, tws, tw.data());
        const int f = c.tid / C::T, j = c.tid % C::T;
        double2* xb = smem + c.grp * C::GROUP_XBUF + f * C::XBUF;
        const double2* src = in.data() + (size_t)(c.grp * C::F + f) * L;
        double2* dst = out.data() + (size_t)(c.grp * C::F + f) * L;
        double2 v[16];
        for (int e = 0; e < 16; ++e) v[e] = src[j + e * C::T];
        fft_regs<L, INV>(v, c, xb, tws, j);
        for (int e = 0; e < 16; ++e) dst[j + e * C::T] = v[e];
    });
    // reference DFT in long double for the first and last transform
    double err = 0, nrm = 0;
    const long double tp = 2.0L * acosl(-1.0L);
    for (int q : {0, nf - 1}) {
        for (int k = 0; k < L; k += (L > 256 ? 37 : 1)) {
            long double sr = 0, si = 0;
            for (int n = 0; n < L; ++n) {
                long double ang = (INV ? tp : -tp) * (long double)(((long long)k * n) % L) / L;
                long double c = cosl(ang), s = sinl(ang);
                sr += in[(size_t)q * L + n].x * c - in[(size_t)q * L + n].y * s;
                si += in[(size_t)q * L + n].x * s + in[(size_t)q * L + n].y * c;
            }
            err = fmax(err, fmax(fabs((double)sr - out[(size_t)q * L + k].x), fabs((double)si - out[(size_t)q * L + k].y)));
            nrm = fmax(nrm, hypot((double)sr, (double)si));
        }
    }
    printf("  fft L=%4d %s  max err %.3e (|X|max %.3e)\n", L, INV ? "inv" : "fwd", err, nrm);
    return err / nrm;
}

// ---------------------------------------------------------------- convolution
// lengths above 4096 go through the radix-2Q bodies of ilm_conv_big.cuh
template <int L> struct Len {
    static constexpr bool big = L > 4096;
    static constexpr int base = big ? 4096 : L;
    static constexpr int Q = big ? L / 4096 : 1;
    using Cfg = FftCfg<base>;
    static constexpr int cpw = big ? 2 : (Cfg::F == 1 ? 2 : Cfg::F);        // pass B / C work item width
    static int rowsA(int nrows) { return big ? nrows : (nrows + Cfg::F - 1) / Cfg::F; }
    static int rowsC(int nrows) { return big ? (Q == 2 ? nrows * Q : nrows) : (nrows + cpw - 1) / cpw; }
};
template <int L> static void runA(HostCtx& c, const ConvArgs& a, double2* sm, int b, int nb) {
    if constexpr (Len<L>::big) passA_big_body<Len<L>::Q>(c, a, sm, b, nb); else passA_body<L>(c, a, sm, b, nb);
}
template <int L, int MODE> static void runB(HostCtx& c, const ConvArgs& a, double2* sm, int b, int nb) {
    if constexpr (Len<L>::big) passB_big_body<Len<L>::Q, MODE>(c, a, sm, b, nb); else passB_body<L, MODE>(c, a, sm, b, nb);
}
// column pass of one emulated CTA (direct lengths) or one emulated cluster (big lengths)
template <int L, int MODE> static void launchB(const ConvArgs& a, int b, int nb) {
    using CY = typename Len<L>::Cfg;
    if constexpr (Len<L>::big) run_cluster(Len<L>::Q, CY::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runB<L, MODE>(c, a, sm, b, nb); });
    else run_cta(CY::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runB<L, MODE>(c, a, sm, b, nb); });
}
template <int L> static void runC(HostCtx& c, const ConvArgs& a, double2* sm, int b, int nb) {
    if constexpr (Len<L>::big && Len<L>::Q == 2) passC_big_percta_body<2>(c, a, sm, b, nb);
    else if constexpr (Len<L>::big) passC_big_body<Len<L>::Q>(c, a, sm, b, nb);
    else passC_body<L>(c, a, sm, b, nb);
}
template <int L> static void launchC(const ConvArgs& a, int b, int nb) {
    using CX = typename Len<L>::Cfg;
    if constexpr (Len<L>::big && Len<L>::Q > 2) run_cluster(Len<L>::Q, CX::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runC<L>(c, a, sm, b, nb); });
    else run_cta(CX::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runC<L>(c, a, sm, b, nb); });
}


template <int LX, int LY>
static double test_conv(int NX, int NY, int mx1, int my1, int mx2, int my2, bool second, int rlo = -1, int rhi = -1,
                        int olo = -1, int ohi = -1) {
    // kernel table g(i,j), 0<=i<NX, 0<=j<NY : something LGF-like
    std::vector<double> G((size_t)NX * NY);
    for (int j = 0; j < NY; ++j)
        for (int i = 0; i < NX; ++i) G[(size_t)j * NX + i] = 0.3 * log(1.0 + i * i + 2.0 * j * j) + 0.1 * cos(0.3 * i) - 0.05 * j;
    using CX = typename Len<LX>::Cfg;
    using CY = typename Len<LY>::Cfg;
    std::vector<double2> twx(CX::TW_TOTAL + 1), twy(CY::TW_TOTAL + 1);
    fft_fill_twiddles<Len<LX>::base>(twx.data(), expm2pii);
    fft_fill_twiddles<Len<LY>::base>(twy.data(), expm2pii);
    ConvArgs a{};
    a.twx = twx.data(); a.twy = twy.data();
    std::vector<double2> wl2(2 * LY), wl2x(2 * LX);
    for (int n = 0; n < 2 * LY; ++n) wl2[n] = expm2pii(n, 2LL * LY);
    for (int n = 0; n < 2 * LX; ++n) wl2x[n] = expm2pii(n, 2LL * LX);
    a.wl2y = wl2.data(); a.wl2x = wl2x.data();
    std::vector<double2> scratch((size_t)3 * 2 * (LY > LX ? LY : LX), cmk(NAN, NAN));        // up to 3 emulated clusters
    a.scratch = scratch.data();
    // ---- Ghat build: h = eps_i eps_j g
    std::vector<double> h((size_t)NX * NY);
    for (int j = 0; j < NY; ++j)
        for (int i = 0; i < NX; ++i) h[(size_t)j * NX + i] = G[(size_t)j * NX + i] * (i ? 2.0 : 1.0) * (j ? 2.0 : 1.0);
    // big LY: rows stored class by class (ConvGeom::rq = Q), MYp a multiple of 2 Q -- what conv_geom() builds on the device
    constexpr int RQ = Len<LY>::big ? Len<LY>::Q : 1;
    auto geom = [&](int MYrows) { const int unit = 2 * RQ; return ConvGeom{LX, LY, MYrows, (MYrows + unit - 1) / unit * unit, RQ}; };
    ConvGeom gg = geom(NY);
    std::vector<double2> S(s_elems(gg)), S2;
    std::vector<double> Ghat(ghat_elems(gg), 0.0);
    a.g = gg; a.f1 = FieldRef{h.data(), NX, NY}; a.f2 = FieldRef{nullptr, 0, 0};
    a.rlo = 0; a.rhi = gg.MYp;
    a.olo = 0; a.ohi = gg.MYp;
    a.S = S.data(); a.GhatOut = Ghat.data(); a.gscale = 1.0 / (4.0 * LX * LY);
    {
        int nwork = Len<LX>::rowsA(gg.MYp);
        int nb = nwork > 3 ? 3 : nwork;     // exercise the persistent loop
        for (int b = 0; b < nb; ++b)
            run_cta(CX::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runA<LX>(c, a, sm, b, nb); });
        nwork = (2 * gg.Lx + Len<LY>::cpw - 1) / Len<LY>::cpw;
        nb = nwork > 3 ? 3 : nwork;
        for (int b = 0; b < nb; ++b)
            launchB<LY, 1>(a, b, nb);
    }
    // ---- fields
    std::mt19937_64 rng(NX * 131 + NY);
    std::normal_distribution<double> nd;
    std::vector<double> w1((size_t)mx1 * my1), w2((size_t)mx2 * my2);
    for (auto& x : w1) x = nd(rng);
    for (auto& x : w2) x = nd(rng);
    if (rlo >= 0) {     // sparse-row input (Schur probe): zero outside [rlo, rhi)
        for (int l = 0; l < my1; ++l) if (l < rlo || l >= rhi) for (int k = 0; k < mx1; ++k) w1[(size_t)l * mx1 + k] = 0;
        for (int l = 0; l < my2; ++l) if (l < rlo || l >= rhi) for (int k = 0; k < mx2; ++k) w2[(size_t)l * mx2 + k] = 0;
    }
    std::vector<double> o1 = w1, o2 = w2;
    if (rlo >= 0) {     // rows outside the range hold garbage that must never be read
        for (int l = 0; l < my1; ++l) if (l < rlo || l >= rhi) for (int k = 0; k < mx1; ++k) o1[(size_t)l * mx1 + k] = NAN;
        for (int l = 0; l < my2; ++l) if (l < rlo || l >= rhi) for (int k = 0; k < mx2; ++k) o2[(size_t)l * mx2 + k] = NAN;
    }
    int MY = second ? (my1 > my2 ? my1 : my2) : my1;
    ConvGeom g2 = geom(MY);
    S.assign(s_elems(g2), cmk(NAN, NAN));
    S2.assign(s_elems(g2), cmk(NAN, NAN));
    a.g = g2; a.f1 = FieldRef{o1.data(), mx1, my1};
    a.rlo = rlo >= 0 ? rlo : 0; a.rhi = rlo >= 0 ? rhi : g2.MYp;
    a.olo = olo >= 0 ? (olo & ~1) : 0; a.ohi = olo >= 0 ? ohi : g2.MYp;     // output rows needed by the caller
    const std::vector<double> in1 = o1, in2 = o2;
    a.f2 = second ? FieldRef{o2.data(), mx2, my2} : FieldRef{nullptr, 0, 0};
    a.S = S.data(); a.S2 = S2.data(); a.Ghat = Ghat.data();
    {
        int nwork = Len<LX>::rowsC(a.ohi - a.olo);
        int nb = nwork > 2 ? 2 : nwork;
        int nworkA = Len<LX>::rowsA(a.rhi - a.rlo);
        int nbA = nworkA > 2 ? 2 : nworkA;
        for (int b = 0; b < nbA; ++b)
            run_cta(CX::SMEM_BYTES, [&](HostCtx& c, double2* sm) { runA<LX>(c, a, sm, b, nbA); });
        int nworkB = (2 * g2.Lx + Len<LY>::cpw - 1) / Len<LY>::cpw;
        int nbB = nworkB > 3 ? 3 : nworkB;
        if constexpr (Len<LY>::big) {
            // the split column pass of the device path: forward sub-transforms -> hand-off block, Q x Q filter in place,
            // inverse half transforms (passB1_big_body / big_filter_point / passB2_big_body<FILTERED>), in two chunks
            constexpr int Q = Len<LY>::Q;
            const int ncols = 2 * g2.Lx, chunk = ncols / 2 + 1;
            std::vector<double2> bigA((size_t)chunk * 2 * LY);
            const unsigned mask = 2u * (unsigned)LY - 1u;
            ConvArgs b2 = a;
            b2.bigA = bigA.data();
            for (int c0 = 0; c0 < ncols; c0 += chunk) {
                b2.bc0 = c0; b2.bnc = std::min(chunk, ncols - c0);
                const int items = b2.bnc * Q, nbs = items > 3 ? 3 : items;
                for (int b = 0; b < nbs; ++b)
                    run_cta(Len<LY>::Cfg::SMEM_BYTES, [&](HostCtx& c, double2* sm) { passB1_big_body<Q>(c, b2, sm, b, nbs); });
                for (int cc = 0; cc < b2.bnc; ++cc)
                    for (int py = 0; py < 2; ++py)
                        for (int kappa = 0; kappa < BIG_M; ++kappa) {
                            const int col = c0 + cc, px = col / g2.Lx, m = col % g2.Lx;
                            double2 wp[Q]; wp[0] = cmk(1.0, 0.0);
                            wp[1] = a.wl2y[(2u * (unsigned)kappa + (unsigned)py) & mask];
                            if constexpr (Q == 4) { wp[2] = cmul(wp[1], wp[1]); wp[3] = cmul(wp[2], wp[1]); }
                            double2* scr = bigA.data() + ((size_t)cc * 2 + py) * (size_t)LY + kappa;
                            const double* gp = a.Ghat + ((size_t)ghat_col(g2, px, m) * 2 + py) * (size_t)LY + kappa;
                            double2 t[Q]; double gh[Q];
                            for (int n1 = 0; n1 < Q; ++n1) { t[n1] = scr[(size_t)n1 * BIG_M]; gh[n1] = gp[(size_t)n1 * BIG_M]; }
                            big_filter_point<Q>(t, gh, wp);
                            for (int n1 = 0; n1 < Q; ++n1) scr[(size_t)n1 * BIG_M] = t[n1];
                        }
                for (int b = 0; b < nbs; ++b)
                    run_cta(Len<LY>::Cfg::SMEM_BYTES, [&](HostCtx& c, double2* sm) { passB2_big_body<Q, true>(c, b2, sm, b, nbs); });
            }
        } else {
            for (int b = 0; b < nbB; ++b)
                launchB<LY, 0>(a, b, nbB);
        }
        for (int b = 0; b < nb; ++b)
            launchC<LX>(a, b, nb);
    }
    // ---- direct check (sampled)
    double err = 0, nrm = 0;
    auto check = [&](const std::vector<double>& w, const std::vector<double>& o, const std::vector<double>& in, int mx, int my) {
        int step = (mx * my > 4000) ? 97 : 1;
        for (int idx = 0; idx < mx * my; idx += step) {
            int i = idx % mx, j = idx / mx;
            if (j < a.olo || j >= a.ohi) {      // rows the caller did not ask for stay untouched
                const bool same = (o[idx] == in[idx]) || (std::isnan(o[idx]) && std::isnan(in[idx]));
                if (!same) err = fmax(err, 1.0);
                continue;
            }
            long double s = 0;
            for (int l = 0; l < my; ++l)
                for (int k = 0; k < mx; ++k) s += (long double)G[(size_t)abs(j - l) * NX + abs(i - k)] * w[(size_t)l * mx + k];
            err = fmax(err, fabs((double)s - o[idx]));
            nrm = fmax(nrm, fabs((double)s));
        }
    };
    check(w1, o1, in1, mx1, my1);
    if (second) check(w2, o2, in2, mx2, my2);
    printf("  conv NX=%d NY=%d Lx=%d Ly=%d fields (%dx%d)%s  max err %.3e (max |out| %.3e)\n", NX, NY, LX, LY, mx1, my1,
           second ? "+2nd" : "", err, nrm);
    return err / nrm;
}

int main() {
    double worst = 0;
    printf("single FFTs\n");
    worst = fmax(worst, test_fft<16, false>());
    worst = fmax(worst, test_fft<32, false>());
    worst = fmax(worst, test_fft<64, true>());
    worst = fmax(worst, test_fft<128, false>());
    worst = fmax(worst, test_fft<256, true>());
    worst = fmax(worst, test_fft<512, false>());
    worst = fmax(worst, test_fft<1024, true>());
    worst = fmax(worst, test_fft<2048, false>());
    worst = fmax(worst, test_fft<4096, false>());
    worst = fmax(worst, test_fft<4096, true>());
    printf("convolutions\n");
    worst = fmax(worst, test_conv<16, 16>(8, 7, 8, 7, 7, 6, true));
    worst = fmax(worst, test_conv<32, 16>(14, 8, 13, 7, 14, 8, true));
    worst = fmax(worst, test_conv<16, 32>(8, 15, 8, 15, 0, 0, false));
    worst = fmax(worst, test_conv<64, 32>(24, 13, 24, 12, 23, 13, true));
    worst = fmax(worst, test_conv<32, 128>(16, 64, 15, 63, 16, 64, true));
    worst = fmax(worst, test_conv<256, 16>(100, 8, 100, 8, 99, 7, true));
    worst = fmax(worst, test_conv<512, 16>(200, 5, 199, 5, 200, 4, true));
    worst = fmax(worst, test_conv<16, 1024>(6, 400, 6, 400, 5, 399, true));
    worst = fmax(worst, test_conv<4096, 16>(1100, 3, 1100, 3, 1099, 2, true));
    worst = fmax(worst, test_conv<16, 4096>(5, 1100, 4, 1100, 5, 1099, true));
    worst = fmax(worst, test_conv<64, 32>(24, 13, 24, 12, 23, 13, true, 5, 9));
    worst = fmax(worst, test_conv<16, 1024>(6, 400, 6, 400, 5, 399, true, 201, 206));
    worst = fmax(worst, test_conv<64, 32>(24, 13, 24, 12, 23, 13, true, 5, 9, 3, 11));      // pruned output rows
    worst = fmax(worst, test_conv<16, 1024>(6, 400, 6, 400, 5, 399, true, 201, 206, 101, 333));
    worst = fmax(worst, test_conv<512, 16>(200, 5, 199, 5, 200, 4, true, 1, 3, 2, 4));
    if (getenv("ILM_TEST_BIG")) {          // lengths above 4096 (radix-2Q step, ilm_conv_big.cuh); slower
        worst = fmax(worst, test_conv<8192, 16>(4500, 3, 4500, 3, 4499, 2, true));
        worst = fmax(worst, test_conv<16, 8192>(3, 4500, 2, 4500, 3, 4499, true));
        worst = fmax(worst, test_conv<16384, 16>(9000, 2, 9000, 2, 8999, 2, true));
        worst = fmax(worst, test_conv<16, 16384>(2, 9000, 2, 9000, 1, 8999, true, 4000, 4007, 3000, 6001));
    }
    printf("worst relative error %.3e -> %s\n", worst, worst < 1e-12 ? "PASS" : "FAIL");
    return worst < 1e-12 ? 0 : 1;
}

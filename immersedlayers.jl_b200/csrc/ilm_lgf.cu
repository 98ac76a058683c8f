// ilm_lgf.cu -- host side of the FFT convolution engine: transform sizes,
// twiddle upload, spectrum buffers, multiplier (Ghat) construction and the
// three-pass apply.  Replaces plan_laplacian / CircularConvolution set-up
// reached from _get_laplacian (src/cache.jl:321-324) and the `L\w` apply
// reached from inverse_laplacian! (src/grid_operators.jl:153-179).
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <vector>

#include "ilm_internal.h"

namespace ilm {

#define ILM_DECL_L(L)                                                                  \
    int conv_launch_L##L(int which, const ConvArgs& a, int nsm, cudaStream_t st, const void* tmap); \
    const double2* conv_twiddles_L##L(size_t* count);
ILM_DECL_L(16) ILM_DECL_L(32) ILM_DECL_L(64) ILM_DECL_L(128) ILM_DECL_L(256)
ILM_DECL_L(512) ILM_DECL_L(1024) ILM_DECL_L(2048) ILM_DECL_L(4096)
int conv_launch_big_Q2(int which, const ConvArgs& a, int nsm, cudaStream_t st, const void* tmap);
int conv_launch_big_Q4(int which, const ConvArgs& a, int nsm, cudaStream_t st, const void* tmap);

conv_launch_fn conv_launcher(int L) {
    switch (L) {
    case 16: return conv_launch_L16;
    case 32: return conv_launch_L32;
    case 64: return conv_launch_L64;
    case 128: return conv_launch_L128;
    case 256: return conv_launch_L256;
    case 512: return conv_launch_L512;
    case 1024: return conv_launch_L1024;
    case 2048: return conv_launch_L2048;
    case 4096: return conv_launch_L4096;
    case 8192: return conv_launch_big_Q2;          // radix-2Q step over the 4096-point transform
    case 16384: return conv_launch_big_Q4;
    default: return nullptr;
    }
}

const double2* conv_twiddles_host(int L, size_t* count) {
    if (L > 4096) L = 4096;                        // big lengths run 4096-point sub-transforms
    switch (L) {
    case 16: return conv_twiddles_L16(count);
    case 32: return conv_twiddles_L32(count);
    case 64: return conv_twiddles_L64(count);
    case 128: return conv_twiddles_L128(count);
    case 256: return conv_twiddles_L256(count);
    case 512: return conv_twiddles_L512(count);
    case 1024: return conv_twiddles_L1024(count);
    case 2048: return conv_twiddles_L2048(count);
    case 4096: return conv_twiddles_L4096(count);
    default: *count = 0; return nullptr;
    }
}

// 3-D tensor map over the S2 spectrum in units of doubles: (8 per 2x2 tile, MYp/2 row pairs,
// Lx tile columns incl. both parities); a box = (4 doubles = one row's two m, 1, <=256 tile
// columns), i.e. the 32-byte pieces of one row gathered into m order (pass C).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_s2_tensor_map(ilm_plan* p, int MYp, const double2* base) {
    if (!base) base = p->S2;
    if (p->tmap_myp == MYp && p->tmap_base == base) return ILM_OK;
    static encode_tiled_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ILM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return ILM_ECUDA; }
        encode = (encode_tiled_fn)fn;
    }
    const int na = p->Lx >> 1;
    cuuint64_t dims[3] = {8, (cuuint64_t)(MYp >> 1), (cuuint64_t)p->Lx};
    cuuint64_t strides[2] = {64, (cuuint64_t)(MYp >> 1) * 64};             // bytes, dims 1 and 2
    cuuint32_t box[3] = {4, 1, (cuuint32_t)(na < 256 ? na : 256)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap*>(p->tmap_s2), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double2*>(base), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed"); return ILM_ECUDA; }
    p->tmap_myp = MYp;
    p->tmap_base = base;
    return ILM_OK;
}

// half padded length: smallest power of two L >= 16 with 2L >= 2n-1
int conv_half_len(int n) {
    int L = 16;
    while (2 * L < 2 * n - 1) L <<= 1;
    return L;
}

int conv_setup(ilm_plan* p) {
    if (const char* e = getenv("ILM_CONV_SKEW_NS")) p->skew_ns = atoi(e);
    if (const char* e = getenv("ILM_PROBE_BAND")) p->band = atoi(e) != 0;
    if (const char* e = getenv("ILM_PROBE_FUSE_E")) p->fuse_e = atoi(e) != 0;
    if (const char* e = getenv("ILM_PROBE_PATCH")) p->patch = atoi(e) != 0;
    if (const char* e = getenv("ILM_PROBE_PRUNE_C")) p->prune_c = atoi(e) != 0;
    if (const char* e = getenv("ILM_SCHUR_SYMM")) p->symm = atoi(e) != 0;
    if (const char* e = getenv("ILM_SCHUR_DUAL")) p->dual = atoi(e) != 0;
    p->Lx = conv_half_len(p->g.NX);
    p->Ly = conv_half_len(p->g.NY);
    if (p->Lx > 16384 || p->Ly > 16384) {
        set_error("grid larger than 16384 cells per direction is not supported by the FFT engine");
        return ILM_ESIZE;
    }
    size_t nx = 0, ny = 0;
    const double2* hx = conv_twiddles_host(p->Lx, &nx);
    const double2* hy = conv_twiddles_host(p->Ly, &ny);
    ILM_CUDA(cudaMalloc(&p->twx, nx * sizeof(double2)));
    ILM_CUDA(cudaMalloc(&p->twy, ny * sizeof(double2)));
    ILM_CUDA(cudaMemcpyAsync(p->twx, hx, nx * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    ILM_CUDA(cudaMemcpyAsync(p->twy, hy, ny * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
    auto upload_wl2 = [&](int L, double2** dst) -> int {           // exp(-2 pi i n / 2L), n < 2L
        std::vector<double2> w(2 * (size_t)L);
        const long double tp = -2.0L * acosl(-1.0L) / (2.0L * L);
        for (size_t n = 0; n < w.size(); ++n) w[n] = cmk((double)cosl(tp * n), (double)sinl(tp * n));
        ILM_CUDA(cudaMalloc(dst, w.size() * sizeof(double2)));
        ILM_CUDA(cudaMemcpy(*dst, w.data(), w.size() * sizeof(double2), cudaMemcpyHostToDevice));
        return ILM_OK;
    };
    ILM_TRY(upload_wl2(p->Ly, &p->wl2y));
    ILM_TRY(upload_wl2(p->Lx, &p->wl2x));
    if (p->Ly > 4096) {
        // split column pass: a chunk of columns whose hand-off block (chunk x 2 Ly complex) stays in L2 between the two launches;
        // nsm columns = 4 (Q = 4) or 2 (Q = 2) items per CTA, 76 MB at L = 16384
        p->big_chunk = p->nsm;
        if (const char* e = getenv("ILM_BIG_CHUNK")) p->big_chunk = atoi(e);
        if (p->big_chunk > 2 * p->Lx) p->big_chunk = 2 * p->Lx;
        if (p->big_chunk > 0) ILM_CUDA(cudaMalloc(&p->bigA, (size_t)p->big_chunk * 2 * (size_t)p->Ly * sizeof(double2)));
    }
    if (p->Ly > 4096 || p->Lx > 4096) {     // hand-off lines of the cluster passes: 2L complex per cluster, <= nsm/2 clusters
        const size_t Lmax = (size_t)(p->Lx > p->Ly ? p->Lx : p->Ly);
        ILM_CUDA(cudaMalloc(&p->conv_scratch, (size_t)p->nsm * Lmax * sizeof(double2)));
    }
    const ConvGeom g = conv_geom(p, p->g.NY);
    p->s_cap = s_elems(g);                    // S / S2 themselves: conv_ensure_spectrum, on first use
    ILM_CUDA(cudaStreamSynchronize(p->stream));
    return ILM_OK;
}

// spectrum geometry of MY field rows: for Ly > 4096 the rows are stored class by class (row interleave Q = Ly / 4096,
// ConvGeom::rq) and MYp is a multiple of 2 Q
ConvGeom conv_geom(const ilm_plan* p, int MY) {
    static const bool no_interleave = getenv("ILM_BIG_NO_INTERLEAVE") != nullptr;
    const int Q = (p->Ly > 4096 && !no_interleave) ? p->Ly / 4096 : 1;
    const int unit = 2 * Q;
    return ConvGeom{p->Lx, p->Ly, MY, (MY + unit - 1) / unit * unit, Q};
}

// the full-size spectrum buffers (2 Lx x MYp complex each), allocated when a full-grid convolution first needs them
int conv_ensure_spectrum(ilm_plan* p, bool need_s2) {
    ilm_plan* owner = p->shared && p->parent ? p->parent : p;
    if (!owner->S) ILM_CUDA(cudaMalloc(&owner->S, owner->s_cap * sizeof(double2)));
    if (need_s2 && !owner->S2) {
        ILM_CUDA(cudaMalloc(&owner->S2, owner->s_cap * sizeof(double2)));
        owner->tmap_myp = -1;
    }
    if (owner != p) { p->S = owner->S; p->S2 = owner->S2; }
    return ILM_OK;
}

void conv_free(ilm_plan* p) {
    cudaFree(p->twx); cudaFree(p->twy); cudaFree(p->wl2y); cudaFree(p->wl2x); cudaFree(p->conv_scratch); cudaFree(p->bigA); cudaFree(p->S); cudaFree(p->S2); cudaFree(p->S2b);
    if (p->stream2) cudaStreamDestroy(p->stream2);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    p->S2b = nullptr; p->stream2 = nullptr; p->ev_fork = p->ev_join = nullptr;
    cudaFree(p->slab_S); cudaFree(p->slab_S2);
    for (auto& k : p->kernels) { cudaFree(k.ghat); cudaFree(k.gxt); }
    cudaFree(p->lgf_dev); p->lgf_dev = nullptr;
    p->kernels.clear();
}

// the plan-constant part of the kernel arguments
ConvArgs conv_base_args(const ilm_plan* p) {
    ConvArgs a{};
    a.S = p->S; a.S2 = p->S2;
    a.twx = p->twx; a.twy = p->twy;
    a.skew_ns = p->skew_ns;
    a.wl2y = p->wl2y; a.wl2x = p->wl2x;
    a.scratch = p->conv_scratch;
    a.bigA = p->bigA; a.bc0 = 0; a.bnc = p->big_chunk;
    return a;
}
static bool use_tma(const ilm_plan* p) { return p->Lx >= 512 && p->Lx <= 4096; }

// table: n x n column-major (host or device), n >= max(NX, NY)
int conv_add_kernel(ilm_plan* p, const double* table, int n, double c0, double factor, int* id) {
    const int NX = p->g.NX, NY = p->g.NY;
    if (n < NX || n < NY) {
        set_error("kernel table smaller than the grid");
        return ILM_ESIZE;
    }
    double* dtab = nullptr;
    ConvKernel k;
    struct Guard {                                  // an error return below frees what this call allocated
        double** tab; ConvKernel* k; bool armed = true;
        ~Guard() { if (armed) { cudaFree(*tab); cudaFree(k->ghat); cudaFree(k->gxt); } }
    } guard{&dtab, &k};
    const double* src = table;
    if (!is_device_ptr(table)) {
        ILM_CUDA(cudaMalloc(&dtab, (size_t)n * NY * sizeof(double)));
        ILM_CUDA(cudaMemcpyAsync(dtab, table, (size_t)n * NY * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        src = dtab;
    }
    // h = eps_i eps_j (G - c0) into scratch field g_a (NX x NY)
    ILM_TRY(conv_ensure_spectrum(p, false));
    ILM_TRY(launch_lgf_prep(p, src, n, NX, NY, c0, p->g_a));
    ConvArgs a = conv_base_args(p);
    a.g = conv_geom(p, NY);
    ILM_CUDA(cudaMalloc(&k.ghat, ghat_elems(a.g) * sizeof(double)));
    a.f1 = FieldRef{p->g_a, NX, NY};
    a.f2 = FieldRef{nullptr, 0, 0};
    a.rlo = 0; a.rhi = a.g.MYp;
    a.olo = 0; a.ohi = a.g.MYp;
    a.GhatOut = k.ghat;
    a.gscale = 1.0 / (4.0 * (double)p->Lx * (double)p->Ly * factor);
    ILM_TRY(conv_launcher(p->Lx)(0, a, p->nsm, p->stream, nullptr));
    if (p->band && p->Lx <= 4096) ILM_TRY(conv_build_gxt(p, a, k, factor));   // x-spectrum of the kernel rows (Schur probes)
    ILM_TRY(conv_launcher(p->Ly)(3, a, p->nsm, p->stream, nullptr));
    p->launches += 2;
    ILM_CUDA(cudaStreamSynchronize(p->stream));
    guard.armed = false;
    if (dtab) {
        if (p->kernels.empty() && !p->lgf_dev) { p->lgf_dev = dtab; p->lgf_ld = n; }   // kernel 0 = LGF: keep the table
        else cudaFree(dtab);
    }
    p->kernels.push_back(k);
    if (id) *id = (int)p->kernels.size() - 1;
    return ILM_OK;
}

bool conv_band_ok(const ilm_plan* p, int kernel_id, int nrows) {
    return kernel_id >= 0 && kernel_id < (int)p->kernels.size() && p->kernels[kernel_id].gxt && p->band && nrows >= 1 &&
           nrows <= conv_band_max_rows();
}

int conv_apply(ilm_plan* p, int kernel_id, FieldRef f1, FieldRef f2, int rlo, int rhi, int olo, int ohi, const ProbeGather* eg) {
    if (kernel_id < 0 || kernel_id >= (int)p->kernels.size()) {
        set_error("unknown convolution kernel id");
        return ILM_EINVAL;
    }
    ILM_TRY(conv_ensure_spectrum(p, true));
    ConvArgs a = conv_base_args(p);
    int MY = f1.p ? f1.my : 0;
    if (f2.p && f2.my > MY) MY = f2.my;
    if (MY == 0) return ILM_OK;
    a.g = conv_geom(p, MY);
    a.f1 = f1; a.f2 = f2;
    a.rlo = rlo < 0 ? 0 : rlo;
    a.rhi = (rhi < 0 || rhi > a.g.MYp) ? a.g.MYp : rhi;
    a.olo = olo < 0 ? 0 : (olo & ~1);
    a.ohi = (ohi < 0 || ohi > a.g.MYp) ? a.g.MYp : ohi;
    if (a.ohi <= a.olo) { a.olo = 0; a.ohi = a.g.MYp < 2 ? a.g.MYp : 2; }
    a.Ghat = p->kernels[kernel_id].ghat;
    if (use_tma(p)) ILM_TRY(make_s2_tensor_map(p, a.g.MYp));
    ILM_TRY(conv_launcher(p->Lx)(0, a, p->nsm, p->stream, nullptr));
    const ConvKernel& k = p->kernels[kernel_id];
    if (eg && !conv_band_ok(p, kernel_id, a.rhi - a.rlo)) { set_error("conv_apply: fused interpolation needs the band pass"); return ILM_EINVAL; }
    if (eg) a.eg = *eg;
    if (conv_band_ok(p, kernel_id, a.rhi - a.rlo)) {
        a.s2_rowmajor = 1;
        ILM_TRY(conv_launch_band(p, a, k));
    } else {
        ILM_TRY(conv_launcher(p->Ly)(1, a, p->nsm, p->stream, nullptr));
    }
    ILM_TRY(conv_launcher(p->Lx)(2, a, p->nsm, p->stream, p->tmap_s2));
    p->launches += 3;
    return ILM_OK;
}

// create_RTLinvR probe from the DDF patches: band pass in patch mode (S is not touched), pass C with the fused interpolation
int conv_apply_patch(ilm_plan* p, int kernel_id, const PatchSrc& ps, int MY, int rlo, int rhi, int olo, int ohi, const ProbeGather& eg,
                     cudaStream_t st, double2* S2alt) {
    if (!st) st = p->stream;
    if (!conv_band_ok(p, kernel_id, rhi - rlo) || !eg.part || !ps.wR || p->Lx > 4096) { set_error("conv_apply_patch: needs the band pass and the fused interpolation"); return ILM_EINVAL; }
    ILM_TRY(conv_ensure_spectrum(p, true));
    ConvArgs a = conv_base_args(p);
    a.g = conv_geom(p, MY);
    a.f1 = FieldRef{nullptr, ps.mx, ps.my};
    a.f2 = FieldRef{nullptr, ps.mx, ps.my};
    a.rlo = rlo; a.rhi = rhi;
    a.olo = olo < 0 ? 0 : (olo & ~1);
    a.ohi = (ohi < 0 || ohi > a.g.MYp) ? a.g.MYp : ohi;
    if (a.ohi <= a.olo) { a.olo = 0; a.ohi = a.g.MYp < 2 ? a.g.MYp : 2; }
    const ConvKernel& k = p->kernels[kernel_id];
    a.Ghat = k.ghat;
    a.eg = eg;
    a.s2_rowmajor = 1;
    if (S2alt) a.S2 = S2alt;
    if (use_tma(p)) ILM_TRY(make_s2_tensor_map(p, a.g.MYp));            // (the row-major path uses plain bulk copies from a.S2)
    ILM_TRY(conv_launch_band(p, a, k, &ps, st));
    ILM_TRY(conv_launcher(p->Lx)(2, a, p->nsm, st, p->tmap_s2));
    p->launches += 2;
    return ILM_OK;
}

// per-pass timing for the roofline report (ilm_profile_conv)
int conv_profile(ilm_plan* p, FieldRef f1, FieldRef f2, int reps, double ms[3], int rlo, int rhi, int olo, int ohi, const ProbeGather* eg,
                 const PatchSrc* ps) {
    ILM_TRY(conv_ensure_spectrum(p, true));
    ConvArgs a = conv_base_args(p);
    int MY = f1.my > f2.my ? f1.my : f2.my;
    a.g = conv_geom(p, MY);
    a.f1 = f1; a.f2 = f2;
    a.rlo = rlo < 0 ? 0 : rlo;
    a.rhi = (rhi < 0 || rhi > a.g.MYp) ? a.g.MYp : rhi;
    a.olo = olo < 0 ? 0 : (olo & ~1);
    a.ohi = (ohi < 0 || ohi > a.g.MYp) ? a.g.MYp : ohi;
    if (a.ohi <= a.olo) { a.olo = 0; a.ohi = a.g.MYp < 2 ? a.g.MYp : 2; }
    a.Ghat = p->kernels[0].ghat;
    cudaEvent_t e0, e1;
    ILM_CUDA(cudaEventCreate(&e0));
    ILM_CUDA(cudaEventCreate(&e1));
    const int Ls[3] = {p->Lx, p->Ly, p->Lx};
    if (use_tma(p)) ILM_TRY(make_s2_tensor_map(p, a.g.MYp));
    const bool band_path = p->kernels[0].gxt && p->band && a.rhi - a.rlo <= conv_band_max_rows();
    a.s2_rowmajor = band_path ? 1 : 0;
    if (eg && band_path) a.eg = *eg;
    for (int which = 0; which < 3; ++which) {
        const ConvKernel& k = p->kernels[0];
        const bool band = which == 1 && band_path;
        const bool patch = band_path && eg && ps && p->patch && p->Lx <= 4096;     // what the create_RTLinvR probes launch
        if (which == 0 && patch) { ms[0] = 0.0; continue; }                        // pass A is not part of that path
        auto launch = [&]() { return band ? conv_launch_band(p, a, k, patch ? ps : nullptr) : conv_launcher(Ls[which])(which, a, p->nsm, p->stream, p->tmap_s2); };
        ILM_TRY(launch());                                                                // warm-up
        ILM_CUDA(cudaEventRecord(e0, p->stream));
        for (int r = 0; r < reps; ++r) ILM_TRY(launch());
        ILM_CUDA(cudaEventRecord(e1, p->stream));
        ILM_CUDA(cudaEventSynchronize(e1));
        float t = 0;
        ILM_CUDA(cudaEventElapsedTime(&t, e0, e1));
        ms[which] = (double)t / reps;
        p->launches += reps + 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ILM_OK;
}

}  // namespace ilm

// ilm_conv_inst.cu -- CUDA kernels of the FFT convolution for one transform
// length, compiled once per -DILM_L=<16..4096> (see Makefile).  512 threads =
// two independent 256-thread groups synchronised with named barriers; one
// persistent CTA per SM.
#include <cuda.h>

#include <cmath>
#include <vector>

#include "ilm_internal.h"

#ifndef ILM_L
#error "compile with -DILM_L=<fft length>"
#endif

namespace ilm {

struct DevCtx {
    int tid, grp;
    const void* tmap;        // CUtensorMap of the S2 spectrum (pass C), else null
    ILM_HD void sync() {
#ifdef __CUDA_ARCH__
        asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
#endif
    }
    ILM_HD void sync_cta() {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    // producer/consumer hand-off between the two groups (512 = both groups)
    ILM_HD void arrive(int id) {
#ifdef __CUDA_ARCH__
        __threadfence_block();
        asm volatile("bar.arrive %0, 512;" ::"r"(id) : "memory");
#endif
    }
    ILM_HD void wait(int id) {
#ifdef __CUDA_ARCH__
        asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory");
#endif
    }
    ILM_HD void delay(int ns) {
#ifdef __CUDA_ARCH__
        if (ns > 0) __nanosleep((unsigned)ns);
#endif
    }
    // ---- bulk-tensor (TMA) row loads of pass C -------------------------------------------
    ILM_HD void tma_init(double2* mbar) {
#ifdef __CUDA_ARCH__
        if (tid == 0) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        sync();
#endif
    }
    // rows row0 .. row0+nrows-1 of parity px (L complex each, m order) into dst + r*slot_stride
    ILM_HD void tma_load_rows(const ConvArgs& a, int px, int row0, int nrows, int L, int slot_stride, double2* dst,
                              double2* mbar) {
#ifdef __CUDA_ARCH__
        if (tid == 0) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
            const unsigned long long tm = reinterpret_cast<unsigned long long>(tmap);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(nrows * L * 16) : "memory");
            const int na = L >> 1;                                 // tile columns per row
            const int box = na < 256 ? na : 256;
            for (int r = 0; r < nrows; ++r) {
                const int row = row0 + r;
                for (int a0 = 0; a0 < na; a0 += box) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * slot_stride + 2 * a0);
                    const int c0 = (row & 1) * 4, c1 = row >> 1, c2 = px * na + a0;
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(d), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(mb) : "memory");
                }
            }
        }
#endif
    }
    ILM_HD void tma_wait(double2* mbar, int phase) {
#ifdef __CUDA_ARCH__
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(mb), "r"(phase) : "memory");
#endif
    }
    ILM_HD void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
    }
};

#define ILM_CAT2(a, b) a##b
#define ILM_CAT(a, b) ILM_CAT2(a, b)

__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passA_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passA_body<ILM_L>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB_body<ILM_L, 0>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passC_L, ILM_L)(ConvArgs a, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(128) double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), &tm};
    passC_body<ILM_L>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passG_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB_body<ILM_L, 1>(c, a, smem, blockIdx.x, gridDim.x);
}

int ILM_CAT(conv_launch_L, ILM_L)(int which, const ConvArgs& a, int nsm, cudaStream_t st, const void* tmap) {
    using C = FftCfg<ILM_L>;
    static bool attr_done_dev[64] = {};           // per device: the attribute belongs to the device's function
    int dev = 0;
    ILM_CUDA(cudaGetDevice(&dev));
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passA_L, ILM_L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passB_L, ILM_L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passC_L, ILM_L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passG_L, ILM_L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done = true;
    }
    int nwork;
    const int cpw = C::F == 1 ? 2 : C::F;          // rows / columns per work item
    if (which == 0) nwork = (a.rhi - a.rlo + C::F - 1) / C::F;
    else if (which == 2) nwork = (a.ohi - a.olo + cpw - 1) / cpw;
    else {
        nwork = (2 * a.g.Lx + cpw - 1) / cpw;
        if (a.whi > 0) nwork = (a.whi < nwork ? a.whi : nwork) - a.wlo;
    }
    int grid = nwork < nsm ? nwork : nsm;
    if (grid < 1) grid = 1;
    switch (which) {
    case 0: ILM_CAT(ilm_passA_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 1: ILM_CAT(ilm_passB_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 2: {
        CUtensorMap tm{};
        if (tmap) tm = *reinterpret_cast<const CUtensorMap*>(tmap);
        ILM_CAT(ilm_passC_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a, tm);
        break;
    }
    default: ILM_CAT(ilm_passG_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    }
    ILM_CUDA(cudaGetLastError());
    return ILM_OK;
}

static double2 expm2pii_ld(long long num, long long den) {
    num %= den;
    long double ang = -2.0L * acosl(-1.0L) * (long double)num / (long double)den;
    return cmk((double)cosl(ang), (double)sinl(ang));
}

const double2* ILM_CAT(conv_twiddles_L, ILM_L)(size_t* count) {
    static std::vector<double2> tw;
    if (tw.empty()) {
        tw.resize(FftCfg<ILM_L>::TW_TOTAL);
        fft_fill_twiddles<ILM_L>(tw.data(), expm2pii_ld);
    }
    *count = tw.size();
    return tw.data();
}

}  // namespace ilm

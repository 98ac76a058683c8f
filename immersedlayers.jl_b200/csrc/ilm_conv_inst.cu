// ilm_conv_inst.cu -- CUDA kernels of the FFT convolution for one transform
// length, compiled once per -DILM_L=<16..4096> (see Makefile).  512 threads =
// two independent 256-thread groups synchronised with named barriers; one
// persistent CTA per SM.
#include <cmath>
#include <vector>

#include "ilm_internal.h"

#ifndef ILM_L
#error "compile with -DILM_L=<fft length>"
#endif

namespace ilm {

struct DevCtx {
    int tid, grp;
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
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8)};
    passA_body<ILM_L>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8)};
    passB_body<ILM_L, 0>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passC_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8)};
    passC_body<ILM_L>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passG_L, ILM_L)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8)};
    passB_body<ILM_L, 1>(c, a, smem, blockIdx.x, gridDim.x);
}

int ILM_CAT(conv_launch_L, ILM_L)(int which, const ConvArgs& a, int nsm, cudaStream_t st) {
    using C = FftCfg<ILM_L>;
    static bool attr_done = false;
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
    else if (which == 2) nwork = (a.g.MYp + cpw - 1) / cpw;
    else nwork = (2 * a.g.Lx + cpw - 1) / cpw;
    int grid = nwork < nsm ? nwork : nsm;
    if (grid < 1) grid = 1;
    switch (which) {
    case 0: ILM_CAT(ilm_passA_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 1: ILM_CAT(ilm_passB_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 2: ILM_CAT(ilm_passC_L, ILM_L)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
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

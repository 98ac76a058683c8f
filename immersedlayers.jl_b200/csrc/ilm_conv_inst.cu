// ilm_conv_inst.cu -- CUDA kernels of the FFT convolution for one transform
// length, compiled once per -DILM_L=<16..4096> (see Makefile).  512 threads =
// two independent 256-thread groups synchronised with named barriers; one
// persistent CTA per SM.
#include <cuda.h>

#include <cmath>
#include <vector>

#include "ilm_internal.h"
#include "ilm_devctx.cuh"

#ifndef ILM_L
#error "compile with -DILM_L=<fft length>"
#endif

namespace ilm {

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

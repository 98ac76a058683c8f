// ilm_conv_big_inst.cu -- CUDA kernels of the FFT convolution for half padded lengths
// L = Q * 4096 (ilm_conv_big.cuh), compiled once per -DILM_Q=<2|4>.  Same CTA shape as the
// direct lengths: 512 threads = two 256-thread groups, one persistent CTA per SM.
#include "ilm_conv_big.cuh"
#include "ilm_devctx.cuh"

#ifndef ILM_Q
#error "compile with -DILM_Q=<2|4>"
#endif

namespace ilm {

#define ILM_CAT2(a, b) a##b
#define ILM_CAT(a, b) ILM_CAT2(a, b)

__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passA_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passA_big_body<ILM_Q>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB_big_body<ILM_Q, 0>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passC_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passC_big_body<ILM_Q>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passG_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB_big_body<ILM_Q, 1>(c, a, smem, blockIdx.x, gridDim.x);
}

int ILM_CAT(conv_launch_big_Q, ILM_Q)(int which, const ConvArgs& a, int nsm, cudaStream_t st, const void*) {
    using C = FftCfg<BIG_M>;
    static bool attr_done_dev[64] = {};
    int dev = 0;
    ILM_CUDA(cudaGetDevice(&dev));
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passA_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passB_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passC_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passG_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done = true;
    }
    int nwork;
    if (which == 0) nwork = a.rhi - a.rlo;              // one row (all of its classes) per work item
    else if (which == 2) nwork = (a.ohi - a.olo) * ILM_Q;
    else {
        nwork = a.g.Lx;                                   // 2-column tiles
        if (a.whi > 0) nwork = (a.whi < nwork ? a.whi : nwork) - a.wlo;
        if (!a.scratch) { set_error("big column pass without a scratch buffer"); return ILM_EINVAL; }
    }
    int grid = nwork < nsm ? nwork : nsm;
    if (grid < 1) grid = 1;
    switch (which) {
    case 0: ILM_CAT(ilm_passA_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 1: ILM_CAT(ilm_passB_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    case 2: ILM_CAT(ilm_passC_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    default: ILM_CAT(ilm_passG_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a); break;
    }
    ILM_CUDA(cudaGetLastError());
    return ILM_OK;
}

}  // namespace ilm

// ilm_conv_big_inst.cu -- CUDA kernels of the FFT convolution for half padded lengths
// L = Q * 4096 (ilm_conv_big.cuh), compiled once per -DILM_Q=<2|4>.  Same CTA shape as the
// direct lengths: 512 threads = two 256-thread groups, one persistent CTA per SM.
#include <cstdlib>

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
    passB_big_body<ILM_Q, 0>(c, a, smem, blockIdx.x / ILM_Q, gridDim.x / ILM_Q);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB1_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB1_big_body<ILM_Q>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB2_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB2_big_body<ILM_Q, false>(c, a, smem, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passB3_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB2_big_body<ILM_Q, true>(c, a, smem, blockIdx.x, gridDim.x);
}
// in-place Q x Q filter of the hand-off block: thread = (column of the chunk, parity, kappa); the same twiddle construction
// as the in-kernel filter (one table entry per 256-thread slot and the compile-time step roots)
__global__ void __launch_bounds__(256) ILM_CAT(k_big_filter_Q, ILM_Q)(ConvArgs a) {
    constexpr int Q = ILM_Q, T = 256;
    const int Lb = a.g.Ly;
    const unsigned mask = 2u * (unsigned)Lb - 1u;
    const int kappa = blockIdx.x * 256 + threadIdx.x;            // < BIG_M
    const int py = blockIdx.y & 1, cc = blockIdx.y >> 1;         // column of the chunk
    if (cc >= a.bnc) return;
    const int c = a.bc0 + cc, px = c / a.g.Lx, m = c % a.g.Lx;
    const int j = kappa & (T - 1), e = kappa / T;
    double2 wp[Q];
    if constexpr (Q == 2) {
        wp[1] = a.wl2y[(2u * (unsigned)kappa + (unsigned)py) & mask];
    } else {
        const double2 wbase = a.wl2y[(2u * (unsigned)j + (unsigned)py) & mask];
        wp[1] = e ? cmul(wbase, wstep<Q>(e)) : wbase;
        wp[2] = cmul(wp[1], wp[1]); wp[3] = cmul(wp[2], wp[1]);
    }
    wp[0] = cmk(1.0, 0.0);
    double2* scr = a.bigA + ((size_t)cc * 2 + py) * (size_t)Lb + kappa;
    const double* gp = a.Ghat + ((size_t)ghat_col(a.g, px, m) * 2 + py) * (size_t)Lb + kappa;
    double2 t[Q];
    double gh[Q];
#pragma unroll
    for (int n1 = 0; n1 < Q; ++n1) { t[n1] = scr[(size_t)n1 * BIG_M]; gh[n1] = gp[(size_t)n1 * BIG_M]; }
    big_filter_point<Q>(t, gh, wp);
#pragma unroll
    for (int n1 = 0; n1 < Q; ++n1) scr[(size_t)n1 * BIG_M] = t[n1];
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passC_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
#if ILM_Q == 2
    passC_big_percta_body<ILM_Q>(c, a, smem, blockIdx.x, gridDim.x);
#else
    passC_big_body<ILM_Q>(c, a, smem, blockIdx.x / ILM_Q, gridDim.x / ILM_Q);
#endif
}
__global__ void __launch_bounds__(512, 1) ILM_CAT(ilm_passG_big_Q, ILM_Q)(ConvArgs a) {
    extern __shared__ double2 smem[];
    DevCtx c{(int)(threadIdx.x & 255), (int)(threadIdx.x >> 8), nullptr};
    passB_big_body<ILM_Q, 1>(c, a, smem, blockIdx.x / ILM_Q, gridDim.x / ILM_Q);
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
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passB1_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passB2_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        ILM_CUDA(cudaFuncSetAttribute(ILM_CAT(ilm_passB3_big_Q, ILM_Q), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done = true;
    }
    if (which == 1 && a.bigA && a.bnc > 0) {
        // split column pass: chunks of bnc columns, three launches per chunk (ilm_conv_big.cuh); a slab owns the tile
        // range [wlo, whi)
        int cbeg = 0, ncols = 2 * a.g.Lx;
        if (a.whi > 0) { cbeg = 2 * a.wlo; ncols = 2 * (a.whi < a.g.Lx ? a.whi : a.g.Lx); }
        ConvArgs b = a;
        for (int c0 = cbeg; c0 < ncols; c0 += a.bnc) {
            b.bc0 = c0;
            b.bnc = ncols - c0 < a.bnc ? ncols - c0 : a.bnc;
            const int items = b.bnc * ILM_Q;
            const int grid = items < nsm ? items : nsm;
            ILM_CAT(ilm_passB1_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(b);
            static const bool fused_filter = getenv("ILM_BIG_FUSED_FILTER") != nullptr;     // the filter inside the inverse kernel (first split form)
            if (fused_filter) {
                ILM_CAT(ilm_passB2_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(b);
            } else {
                ILM_CAT(k_big_filter_Q, ILM_Q)<<<dim3(BIG_M / 256, 2 * b.bnc), 256, 0, st>>>(b);
                ILM_CAT(ilm_passB3_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(b);
            }
        }
        ILM_CUDA(cudaGetLastError());
        return ILM_OK;
    }
    if (which == 0 || (which == 2 && ILM_Q == 2)) {
        const int nwork = which == 0 ? a.rhi - a.rlo : (a.ohi - a.olo) * ILM_Q;
        int grid = nwork < nsm ? nwork : nsm;
        if (grid < 1) grid = 1;
        if (which == 0) ILM_CAT(ilm_passA_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a);
        else ILM_CAT(ilm_passC_big_Q, ILM_Q)<<<grid, 512, C::SMEM_BYTES, st>>>(a);
    } else {
        // column pass / last row pass: one cluster of ILM_Q CTAs per 2-column tile / per row, as many clusters
        // as fit the device; the cluster's hand-off line lives in a.scratch
        if (!a.scratch) { set_error("big-length pass without a scratch buffer"); return ILM_EINVAL; }
        int nitems = which == 2 ? a.ohi - a.olo : a.g.Lx;
        if (which != 2 && a.whi > 0) nitems = (a.whi < nitems ? a.whi : nitems) - a.wlo;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ILM_Q; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        static int max_clusters[64] = {};
        if (!max_clusters[dev & 63]) {
            cfg.gridDim = dim3((nsm / ILM_Q) * ILM_Q);
            int n = 0;
            ILM_CUDA(cudaOccupancyMaxActiveClusters(&n, ILM_CAT(ilm_passB_big_Q, ILM_Q), &cfg));
            max_clusters[dev & 63] = n > 0 ? (n < nsm / ILM_Q ? n : nsm / ILM_Q) : 1;
        }
        int nclusters = nitems < max_clusters[dev & 63] ? nitems : max_clusters[dev & 63];
        if (nclusters < 1) nclusters = 1;
        cfg.gridDim = dim3(nclusters * ILM_Q);
        if (which == 1) ILM_CUDA(cudaLaunchKernelEx(&cfg, ILM_CAT(ilm_passB_big_Q, ILM_Q), a));
        else if (which == 2) ILM_CUDA(cudaLaunchKernelEx(&cfg, ILM_CAT(ilm_passC_big_Q, ILM_Q), a));
        else ILM_CUDA(cudaLaunchKernelEx(&cfg, ILM_CAT(ilm_passG_big_Q, ILM_Q), a));
    }
    ILM_CUDA(cudaGetLastError());
    return ILM_OK;
}

}  // namespace ilm

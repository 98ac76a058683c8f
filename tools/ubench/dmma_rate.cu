// FP64 tensor (DMMA m8n8k4) issue-rate probe for B200 (tools/ubench, not product): independent accumulators per warp,
// 1..16 warps per SM, against the DFMA rate of fp64rate.cu.  Feeds the trailing update of the LU (ilm_dense.cu).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC> __global__ void k(int iters, double* out) {
    double acc[NACC][2];
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC> void run(int threads) {
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, nsm = 148;
    k<NACC><<<nsm, threads>>>(10, out);
    cudaEventRecord(e0);
    k<NACC><<<nsm, threads>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 8 * 4 * (double)NACC * iters * (threads / 32) * nsm;
    printf("warps/SM %2d  independent accumulators %2d : %7.2f TFLOP/s FP64 tensor\n", threads / 32, NACC, flop / (ms * 1e-3) * 1e-12);
    cudaFree(out);
}

int main() {
    for (int threads : {32, 128, 256, 512}) {
        run<1>(threads);
        run<4>(threads);
        run<16>(threads);
    }
    return 0;
}

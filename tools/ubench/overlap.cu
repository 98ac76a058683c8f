// Micro-benchmark: can FP64 butterflies of one 256-thread group overlap with the
// shared-memory exchange of the other group on a B200 SM?  (tools/ubench, not product)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void compute_phase(double (&r)[32], double c, int n) {
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = fma(r[i], c, r[(i + 7) & 31]);   // 32 independent-ish DFMA
    }
}
__device__ __forceinline__ void smem_phase(double2* sm, double2 (&v)[16], int tid, int n) {
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int t = 0; t < 16; ++t) sm[(tid * 16 + t) + ((tid * 16 + t) >> 4)] = v[t];
        asm volatile("bar.sync %0, 256;" ::"r"(1 + (int)(threadIdx.x >> 8)) : "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = sm[(tid + e * 256) + ((tid + e * 256) >> 4)];
        asm volatile("bar.sync %0, 256;" ::"r"(1 + (int)(threadIdx.x >> 8)) : "memory");
    }
}

// mode 0: both groups compute only; 1: both smem only; 2: group0 compute, group1 smem;
// 3: both alternate compute/smem in phase; 4: alternate, group 1 starts with smem (anti-phase)
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, int ncomp, double* out) {
    extern __shared__ double2 smem[];
    const int grp = threadIdx.x >> 8, tid = threadIdx.x & 255;
    double2* sm = smem + grp * 4352;
    double r[32];
    double2 v[16];
    for (int i = 0; i < 32; ++i) r[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int i = 0; i < 16; ++i) v[i] = make_double2(r[i], r[i + 16]);
    const double c = 0.999999;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) compute_phase(r, c, ncomp);
        else if (mode == 1) smem_phase(sm, v, tid, 1);
        else if (mode == 2) { if (grp == 0) compute_phase(r, c, ncomp); else smem_phase(sm, v, tid, 1); }
        else if (mode == 5) { if (grp == 0) compute_phase(r, c, ncomp); }
        else if (mode == 6) { if (grp == 1) smem_phase(sm, v, tid, 1); }
        else if (mode == 3) { compute_phase(r, c, ncomp); smem_phase(sm, v, tid, 1); }
        else {
            if (grp == 0) { compute_phase(r, c, ncomp); smem_phase(sm, v, tid, 1); }
            else { smem_phase(sm, v, tid, 1); compute_phase(r, c, ncomp); }
        }
    }
    double s = 0;
    for (int i = 0; i < 32; ++i) s += r[i];
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4352 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int ncomp : {2, 4, 8, 16})
    for (int mode : {5, 6, 2}) {
        printf("ncomp %2d ", ncomp);
        k<<<148, 512, 2 * 4352 * 16>>>(mode, 10, ncomp, out);
        cudaEventRecord(e0);
        k<<<148, 512, 2 * 4352 * 16>>>(mode, iters, ncomp, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d: %.3f ms  -> %.0f cycles/iter at %.2f GHz (nominal max clock)\n", mode, ms, ms * 1e-3 * clk * 1e3 / iters, clk * 1e-6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// FP64 issue-rate probe for B200 (tools/ubench, not product): DADD / DMUL / DFMA / mixed
// chains with 32 independent accumulators per thread, 8 or 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __global__ void __launch_bounds__(512, 1) k(int iters, double c, double* out) {
    double r[32];
    for (int i = 0; i < 32; ++i) r[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (OP == 0) r[i] = __dadd_rn(r[i], c);
            else if (OP == 1) r[i] = __dmul_rn(r[i], c);
            else if (OP == 2) r[i] = fma(r[i], c, c);
            else if (OP == 3) r[i] = (i & 1) ? __dadd_rn(r[i], c) : __dmul_rn(r[i], c);
            else r[i] = (i & 1) ? __dadd_rn(r[i], c) : fma(r[i], c, c);
        }
    }
    double s = 0;
    for (int i = 0; i < 32; ++i) s += r[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP> void run(const char* name, int threads) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<OP><<<148, threads>>>(10, 0.9999999, out);
    cudaEventRecord(e0);
    k<OP><<<148, threads>>>(iters, 0.9999999, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)iters * 32 * (threads / 32) / 4;      // warp-instr per SMSP
    printf("%-10s %3d thr: %.3f ms, %.2f ns per warp-instr per SMSP (2 cyc @1.965GHz = 1.02 ns)\n", name, threads, ms, ms * 1e6 / winst);
    cudaFree(out);
}

int main() {
    for (int th : {256, 512}) {
        run<0>("DADD", th); run<1>("DMUL", th); run<2>("DFMA", th); run<3>("DADD+DMUL", th); run<4>("DADD+DFMA", th);
    }
    return 0;
}

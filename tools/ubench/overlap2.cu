// Which memory-side activity of one 256-thread group slows the FP64 stream of the other group
// on a B200 SM?  (tools/ubench, not product)
#include <cstdio>
#include <cuda_runtime.h>

template <bool F64> __device__ __forceinline__ void compute_phase(double (&r)[32], float (&q)[32], int n) {
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (F64) r[i] = fma(r[i], 0.999999, 1e-9);
            else q[i] = fmaf(q[i], 0.999999f, 1e-9f);
        }
    }
}

template <int BYTES> __device__ __forceinline__ void smem_phase(char* smc, int tid, double2 (&v)[16]) {
    const int bar = 1 + (int)(threadIdx.x >> 8);
    if (BYTES == 16) {
        double2* sm = (double2*)smc;
#pragma unroll
        for (int t = 0; t < 16; ++t) sm[17 * tid + t] = v[t];
        asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = sm[(tid + e * 256) + ((tid + e * 256) >> 4)];
        asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
    } else if (BYTES == 8) {          // same bytes as two 8-byte planes
        double* sm = (double*)smc;
#pragma unroll
        for (int t = 0; t < 16; ++t) { sm[17 * tid + t] = v[t].x; sm[4352 + 17 * tid + t] = v[t].y; }
        asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) { int i = (tid + e * 256) + ((tid + e * 256) >> 4); v[e].x = sm[i]; v[e].y = sm[4352 + i]; }
        asm volatile("bar.sync %0, 256;" ::"r"(bar) : "memory");
    }
}

// mode: 0 fp64 only (grp0) ; 1 smem16 only (grp1); 2 fp64 || smem16 ; 3 fp32 only; 4 fp32 || smem16 ;
//       5 smem8 only; 6 fp64 || smem8 ; 7 global-load only (grp1); 8 fp64 || global-load
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, const double2* __restrict__ gsrc, double* out) {
    extern __shared__ char smem[];
    const int grp = threadIdx.x >> 8, tid = threadIdx.x & 255;
    char* sm = smem + grp * 4352 * 16;
    double r[32]; float q[32]; double2 v[16];
    for (int i = 0; i < 32; ++i) { r[i] = 1.0 + 1e-9 * (threadIdx.x + i); q[i] = (float)r[i]; }
    for (int i = 0; i < 16; ++i) v[i] = make_double2(r[i], r[i + 16]);
    const bool do_f64 = (mode == 0 || mode == 2 || mode == 6 || mode == 8) && grp == 0;
    const bool do_f32 = (mode == 3 || mode == 4) && grp == 0;
    const bool do_s16 = (mode == 1 || mode == 2 || mode == 4) && grp == 1;
    const bool do_s8 = (mode == 5 || mode == 6) && grp == 1;
    const bool do_g = (mode == 7 || mode == 8) && grp == 1;
    for (int it = 0; it < iters; ++it) {
        if (do_f64) compute_phase<true>(r, q, 8);
        if (do_f32) compute_phase<false>(r, q, 8);
        if (do_s16) smem_phase<16>(sm, tid, v);
        if (do_s8) smem_phase<8>(sm, tid, v);
        if (do_g) {
            const double2* p = gsrc + ((size_t)blockIdx.x * 64 + (it & 63)) * 4096 + tid;
#pragma unroll
            for (int e = 0; e < 16; ++e) { double2 t = p[e * 256]; v[e].x += t.x; v[e].y += t.y; }
        }
    }
    double s = 0;
    for (int i = 0; i < 32; ++i) s += r[i] + q[i];
    for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    double2* g; cudaMalloc(&g, (size_t)148 * 64 * 4096 * 16); cudaMemset(g, 0, (size_t)148 * 64 * 4096 * 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4352 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"fp64 alone", "smem.128 alone", "fp64 || smem.128", "fp32 alone", "fp32 || smem.128",
                           "smem.64 alone", "fp64 || smem.64", "LDG.128 alone", "fp64 || LDG.128"};
    const int iters = 2000;
    for (int mode = 0; mode < 9; ++mode) {
        k<<<148, 512, 2 * 4352 * 16>>>(mode, 10, g, out);
        cudaEventRecord(e0);
        k<<<148, 512, 2 * 4352 * 16>>>(mode, iters, g, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-20s %.3f ms -> %5.0f cycles/iter\n", names[mode], ms, ms * 1e-3 * 1.965e9 / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

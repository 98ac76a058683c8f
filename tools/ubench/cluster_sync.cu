// Cluster barrier / DSMEM exchange latency probe for B200 (tools/ubench, not product).  Feeds the design of the LU
// panel (ilm_dense.cu: one DSMEM push + one cluster barrier per column) and of the cluster triangular solve.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cluster_sync cluster_sync.cu && ./cluster_sync
// Prints microseconds per iteration of: __syncthreads, cluster.sync(), DSMEM push of 32 doubles to every peer +
// cluster.sync(), for cluster sizes 1..16 and 256 / 512 / 1024 threads per CTA.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE> __global__ void k(int iters, double* out) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double box[2][16][32];
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            __syncthreads();
        } else if (MODE == 1) {
            cl.sync();
        } else {
            if (wid < C) cl.map_shared_rank(&box[it & 1][rank][0], wid)[lane] = acc + it;
            cl.sync();
            for (int p = 0; p < C; ++p) acc += box[it & 1][p][lane];
        }
    }
    if (acc == 12345.678) out[0] = acc;
    cl.sync();
}

template <int MODE> void run(const char* name, int C, int threads) {
    double* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (cudaLaunchKernelEx(&cfg, k<MODE>, 10, out) != cudaSuccess) { printf("%-28s C=%2d threads=%4d  launch failed: %s\n", name, C, threads, cudaGetErrorString(cudaGetLastError())); return; }
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, k<MODE>, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s C=%2d threads=%4d  %.3f us / iteration\n", name, C, threads, ms * 1e3 / iters);
    cudaFree(out);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("__syncthreads", 1, threads);
        for (int C : {1, 2, 4, 8, 16}) run<1>("cluster.sync", C, threads);
        for (int C : {2, 4, 8, 16}) run<2>("DSMEM push + cluster.sync", C, threads);
    }
    return 0;
}

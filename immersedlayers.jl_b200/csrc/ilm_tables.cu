// ilm_tables.cu -- DDF window tables and cell-bucketed gather lists.
// Compiled with --fmad=false (see ilm_ddf.h): the weights are bit-exact with
// the oracle.  Replaces _get_regularization / _regularization_matrix /
// _interpolation_matrix (src/cache.jl:305-347): instead of N functor
// applications producing CSC matrices, one thread per surface point evaluates
// its <=4x4 window, and the host builds the per-cell gather list (sorted by
// cell, then by point index) that makes regularization an atomics-free gather.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ilm_ddf.h"
#include "ilm_internal.h"

namespace ilm {

__global__ void k_point_tables(int N, const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ ds, double dx, int I0x, int I0y, int ddf, int symmetric,
                               double sx, double sy, int mx, int my, int* __restrict__ i0, int* __restrict__ j0,
                               double* __restrict__ wR, double* __restrict__ wE) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double xs = x[k] / dx + (double)I0x;
    const double ys = y[k] / dx + (double)I0y;
    const double wgt = symmetric ? 1.0 : ds[k] / (dx * dx);
    const int W2 = ddf_width(ddf) * ddf_width(ddf);
    double r[16], e[16];
    int a, b;
    ddf_point_table(ddf, xs, ys, wgt, symmetric != 0, sx, sy, mx, my, &a, &b, r, e);
    i0[k] = a;
    j0[k] = b;
    for (int s = 0; s < W2; ++s) {
        wR[(size_t)k * W2 + s] = r[s];
        wE[(size_t)k * W2 + s] = e[s];
    }
}

void free_table(DevTable& t) {
    cudaFree(t.i0); cudaFree(t.j0); cudaFree(t.wR); cudaFree(t.wE);
    cudaFree(t.cell_idx); cudaFree(t.cell_off); cudaFree(t.ent); cudaFree(t.rowsum);
    t = DevTable();
}

int build_tables(ilm_plan* p) {
    const int N = p->N;
    const int W = ddf_width(p->ddf), W2 = W * W;
    for (int layout = 0; layout < 4; ++layout) {
        DevTable& t = p->tab[layout];
        free_table(t);
        const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
        t.W = W; t.mx = li.mx; t.my = li.my;
        const size_t np = (size_t)(N > 0 ? N : 1);
        ILM_CUDA(cudaMalloc(&t.i0, np * sizeof(int)));
        ILM_CUDA(cudaMalloc(&t.j0, np * sizeof(int)));
        ILM_CUDA(cudaMalloc(&t.wR, np * W2 * sizeof(double)));
        ILM_CUDA(cudaMalloc(&t.wE, np * W2 * sizeof(double)));
        std::vector<int> hi(N), hj(N);
        std::vector<double> hw((size_t)N * W2);
        if (N > 0) {
            k_point_tables<<<(N + 127) / 128, 128, 0, p->stream>>>(N, p->x, p->y, p->ds, p->g.dx, p->g.I0x, p->g.I0y,
                                                                    p->ddf, p->scaling == ILM_INDEX_SCALING, li.sx,
                                                                    li.sy, li.mx, li.my, t.i0, t.j0, t.wR, t.wE);
            ILM_CUDA(cudaGetLastError());
            p->launches++;
            ILM_CUDA(cudaMemcpyAsync(hi.data(), t.i0, N * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
            ILM_CUDA(cudaMemcpyAsync(hj.data(), t.j0, N * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
            ILM_CUDA(cudaStreamSynchronize(p->stream));
        }
        t.h_j0 = hj;
        // gather list: (cell, k, slot) for every in-range window entry, sorted by cell then k
        struct Ent { int cell, id; };
        std::vector<Ent> ents;
        ents.reserve((size_t)N * W2);
        for (int k = 0; k < N; ++k)
            for (int b = 0; b < W; ++b)
                for (int a = 0; a < W; ++a) {
                    const int i = hi[k] + a, j = hj[k] + b;
                    if (i >= 0 && i < li.mx && j >= 0 && j < li.my) ents.push_back({i + li.mx * j, k * W2 + b * W + a});
                }
        std::stable_sort(ents.begin(), ents.end(), [](const Ent& u, const Ent& v) { return u.cell < v.cell; });
        std::vector<int> cell_idx, cell_off, ent(ents.size());
        for (size_t q = 0; q < ents.size(); ++q) {
            if (q == 0 || ents[q].cell != ents[q - 1].cell) {
                cell_idx.push_back(ents[q].cell);
                cell_off.push_back((int)q);
            }
            ent[q] = ents[q].id;
        }
        cell_off.push_back((int)ents.size());
        t.ncell = (int)cell_idx.size();
        t.nent = (int)ents.size();
        ILM_CUDA(cudaMalloc(&t.cell_idx, (cell_idx.size() + 1) * sizeof(int)));
        ILM_CUDA(cudaMalloc(&t.cell_off, cell_off.size() * sizeof(int)));
        ILM_CUDA(cudaMalloc(&t.ent, (ent.size() + 1) * sizeof(int)));
        ILM_CUDA(cudaMalloc(&t.rowsum, (cell_idx.size() + 1) * sizeof(double)));
        if (t.ncell) ILM_CUDA(cudaMemcpyAsync(t.cell_idx, cell_idx.data(), cell_idx.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        ILM_CUDA(cudaMemcpyAsync(t.cell_off, cell_off.data(), cell_off.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        if (t.nent) ILM_CUDA(cudaMemcpyAsync(t.ent, ent.data(), ent.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        ILM_CUDA(cudaStreamSynchronize(p->stream));
        if (layout == ILM_NODES_PRIMAL) ILM_TRY(launch_filter_rowsum(p, t));
    }
    return ILM_OK;
}

}  // namespace ilm

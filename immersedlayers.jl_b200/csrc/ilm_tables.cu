// ilm_tables.cu -- DDF window tables and cell-bucketed gather lists.
// Compiled with --fmad=false (see ilm_ddf.h): the weights are bit-exact with
// the oracle.  Replaces _get_regularization / _regularization_matrix /
// _interpolation_matrix (src/cache.jl:305-347): instead of N functor
// applications producing CSC matrices, one thread per surface point evaluates
// its <=4x4 window, and the host builds the per-cell gather list (sorted by
// cell, then by point index) that makes regularization an atomics-free gather.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
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
    cudaFree(t.row_ptr); cudaFree(t.row_ent); cudaFree(t.part); cudaFree(t.need);
    t = DevTable();
}

// Device buffers of the tables are grow-only: update_points on a moving body (src/system.jl:26-50) rebuilds the tables
// every step, and a cudaFree / cudaMalloc per buffer and refresh would synchronise the device 64 times.

// stable LSD radix sort of (cell, id) pairs by cell (3 passes of 10 bits cover 2^30 > 16384^2 cells); entries are
// generated in ascending point order, so ties keep that order -- the order of a CSC column scan, as std::stable_sort gave
static void sort_by_cell(std::vector<int>& cell, std::vector<int>& id, std::vector<int>& cell_tmp, std::vector<int>& id_tmp) {
    const size_t n = cell.size();
    cell_tmp.resize(n);
    id_tmp.resize(n);
    int maxc = 0;
    for (size_t q = 0; q < n; ++q) maxc = std::max(maxc, cell[q]);
    for (int shift = 0; shift < 30 && (maxc >> shift) > 0; shift += 10) {
        size_t count[1025] = {0};
        for (size_t q = 0; q < n; ++q) count[((cell[q] >> shift) & 1023) + 1]++;
        for (int b = 0; b < 1024; ++b) count[b + 1] += count[b];
        for (size_t q = 0; q < n; ++q) {
            const size_t dst = count[(cell[q] >> shift) & 1023]++;
            cell_tmp[dst] = cell[q];
            id_tmp[dst] = id[q];
        }
        cell.swap(cell_tmp);
        id.swap(id_tmp);
    }
}

int build_tables(ilm_plan* p) {
    static const bool ttrace = getenv("ILM_TABLE_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
    const auto t_begin = now();
    const int N = p->N;
    const int W = ddf_width(p->ddf), W2 = W * W;
    const size_t np = (size_t)(N > 0 ? N : 1);
    std::vector<int> hi[4], hj[4];
    // phase 1: the window tables of all four layouts, one synchronisation
    for (int layout = 0; layout < 4; ++layout) {
        DevTable& t = p->tab[layout];
        const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
        t.W = W; t.mx = li.mx; t.my = li.my;
        if (np * W2 > t.cap_pts || !t.i0) {
            cudaFree(t.i0); cudaFree(t.j0); cudaFree(t.wR); cudaFree(t.wE);
            t.i0 = t.j0 = nullptr; t.wR = t.wE = nullptr;
            const size_t cap = np + np / 4 + 16;
            ILM_CUDA(cudaMalloc(&t.i0, cap * sizeof(int)));
            ILM_CUDA(cudaMalloc(&t.j0, cap * sizeof(int)));
            ILM_CUDA(cudaMalloc(&t.wR, cap * W2 * sizeof(double)));
            ILM_CUDA(cudaMalloc(&t.wE, cap * W2 * sizeof(double)));
            t.cap_pts = cap * W2;
        }
        hi[layout].resize(N);
        hj[layout].resize(N);
        if (N > 0) {
            k_point_tables<<<(N + 127) / 128, 128, 0, p->stream>>>(N, p->x, p->y, p->ds, p->g.dx, p->g.I0x, p->g.I0y,
                                                                    p->ddf, p->scaling == ILM_INDEX_SCALING, li.sx,
                                                                    li.sy, li.mx, li.my, t.i0, t.j0, t.wR, t.wE);
            ILM_CUDA(cudaGetLastError());
            p->launches++;
            ILM_CUDA(cudaMemcpyAsync(hi[layout].data(), t.i0, N * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
            ILM_CUDA(cudaMemcpyAsync(hj[layout].data(), t.j0, N * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        }
    }
    ILM_CUDA(cudaStreamSynchronize(p->stream));
    const double t_phase1 = ms_since(t_begin);
    // phase 2: gather lists (cell, k, slot) for every in-range window entry, sorted by cell then k.  Pure host work per
    // layout (entry generation, radix sort, run detection): the four layouts are built by four host threads, the uploads
    // follow in layout order.
    struct HostLists { std::vector<int> id, cell_idx, cell_off, rptr, rent; std::vector<unsigned short> need; };
    HostLists lists[4];
    auto build_layout = [&](int layout) {
        HostLists& h = lists[layout];
        const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
        const std::vector<int>& xi = hi[layout];
        const std::vector<int>& yj = hj[layout];
        std::vector<int> cell, cell_tmp, id_tmp;
        cell.reserve((size_t)N * W2);
        h.id.reserve((size_t)N * W2);
        for (int k = 0; k < N; ++k)
            for (int b = 0; b < W; ++b)
                for (int a = 0; a < W; ++a) {
                    const int i = xi[k] + a, j = yj[k] + b;
                    if (i >= 0 && i < li.mx && j >= 0 && j < li.my) { cell.push_back(i + li.mx * j); h.id.push_back(k * W2 + b * W + a); }
                }
        sort_by_cell(cell, h.id, cell_tmp, id_tmp);
        for (size_t q = 0; q < cell.size(); ++q)
            if (q == 0 || cell[q] != cell[q - 1]) {
                h.cell_idx.push_back(cell[q]);
                h.cell_off.push_back((int)q);
            }
        h.cell_off.push_back((int)cell.size());
        if (layout == ILM_NODES_PRIMAL) {
            // row buckets of the window rows (k*W + b sorted by grid row, then point) for the fused interpolation
            // of the Schur probes in pass C (ConvArgs::eg)
            h.rptr.assign((size_t)li.my + 1, 0);
            for (int k = 0; k < N; ++k)
                for (int b = 0; b < W; ++b) { const int j = yj[k] + b; if (j >= 0 && j < li.my) ++h.rptr[j + 1]; }
            for (int j = 0; j < li.my; ++j) h.rptr[j + 1] += h.rptr[j];
            h.rent.resize((size_t)h.rptr[li.my]);
            std::vector<int> fill(h.rptr.begin(), h.rptr.end() - 1);
            for (int k = 0; k < N; ++k)
                for (int b = 0; b < W; ++b) { const int j = yj[k] + b; if (j >= 0 && j < li.my) h.rent[fill[j]++] = k * W + b; }
            // columns of every row that some window reads, as one 16-bit word per transform thread: in pass C thread j of a
            // length-L inverse transform ends up with the columns j + e*T (T = L/16), bit e of need[row*T + j] says whether
            // column j + e*T is read.  Threads (warps) without a needed column skip the last butterfly pass and the hand-off.
            const int T = p->Lx / 16;
            h.need.assign((size_t)li.my * T, 0);
            for (int k = 0; k < N; ++k)
                for (int b = 0; b < W; ++b) {
                    const int j = yj[k] + b;
                    if (j < 0 || j >= li.my) continue;
                    for (int a = 0; a < W; ++a) {
                        const int i = xi[k] + a;
                        if (i >= 0 && i < li.mx) h.need[(size_t)j * T + i % T] |= (unsigned short)(1u << (i / T));
                    }
                }
        }
    };
    {
        std::thread workers[3];
        for (int layout = 1; layout < 4; ++layout) workers[layout - 1] = std::thread(build_layout, layout);
        build_layout(0);
        for (auto& w : workers) w.join();
    }
    const double t_phase2 = ms_since(t_begin);
    for (int layout = 0; layout < 4; ++layout) {
        DevTable& t = p->tab[layout];
        HostLists& h = lists[layout];
        const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
        t.h_j0 = hj[layout];
        t.ncell = (int)h.cell_idx.size();
        t.nent = (int)h.id.size();
        if ((size_t)t.ncell + 1 > t.cap_cells || !t.cell_idx) {
            cudaFree(t.cell_idx); cudaFree(t.cell_off); cudaFree(t.rowsum);
            t.cell_idx = t.cell_off = nullptr; t.rowsum = nullptr;
            const size_t cap = (size_t)t.ncell + t.ncell / 4 + 16;
            ILM_CUDA(cudaMalloc(&t.cell_idx, cap * sizeof(int)));
            ILM_CUDA(cudaMalloc(&t.cell_off, (cap + 1) * sizeof(int)));
            ILM_CUDA(cudaMalloc(&t.rowsum, cap * sizeof(double)));
            t.cap_cells = cap;
        }
        if ((size_t)t.nent + 1 > t.cap_ents || !t.ent) {
            cudaFree(t.ent);
            t.ent = nullptr;
            const size_t cap = (size_t)t.nent + t.nent / 4 + 16;
            ILM_CUDA(cudaMalloc(&t.ent, cap * sizeof(int)));
            t.cap_ents = cap;
        }
        // pageable sources: the vectors stay alive until the synchronisation below
        if (t.ncell) ILM_CUDA(cudaMemcpyAsync(t.cell_idx, h.cell_idx.data(), h.cell_idx.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        ILM_CUDA(cudaMemcpyAsync(t.cell_off, h.cell_off.data(), h.cell_off.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        if (t.nent) ILM_CUDA(cudaMemcpyAsync(t.ent, h.id.data(), h.id.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        if (layout == ILM_NODES_PRIMAL) {
            ILM_TRY(launch_filter_rowsum(p, t));
            if ((size_t)li.my + 1 > t.cap_rows || !t.row_ptr) {
                cudaFree(t.row_ptr); t.row_ptr = nullptr;
                ILM_CUDA(cudaMalloc(&t.row_ptr, ((size_t)li.my + 1) * sizeof(int)));
                t.cap_rows = (size_t)li.my + 1;
            }
            if (np * W > t.cap_rowent || !t.row_ent) {
                cudaFree(t.row_ent); cudaFree(t.part); t.row_ent = nullptr; t.part = nullptr;
                const size_t cap = (np + np / 4 + 16) * W;
                ILM_CUDA(cudaMalloc(&t.row_ent, cap * sizeof(int)));
                ILM_CUDA(cudaMalloc(&t.part, cap * PART_BATCH * sizeof(double2)));
                t.part_stride = cap;
                t.cap_rowent = cap;
            }
            if (h.need.size() > t.cap_need || !t.need) {
                cudaFree(t.need); t.need = nullptr;
                ILM_CUDA(cudaMalloc(&t.need, h.need.size() * sizeof(unsigned short) + 16));
                t.cap_need = h.need.size();
            }
            t.need_T = p->Lx / 16;
            if (!h.need.empty()) ILM_CUDA(cudaMemcpyAsync(t.need, h.need.data(), h.need.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, p->stream));
            ILM_CUDA(cudaMemcpyAsync(t.row_ptr, h.rptr.data(), h.rptr.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
            if (!h.rent.empty()) ILM_CUDA(cudaMemcpyAsync(t.row_ent, h.rent.data(), h.rent.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        }
    }
    ILM_CUDA(cudaStreamSynchronize(p->stream));
    if (ttrace) fprintf(stderr, "ilm table refresh, N = %d: window tables + D2H %.3f ms, host lists %.3f ms, uploads %.3f ms\n", N, t_phase1,
                        t_phase2 - t_phase1, ms_since(t_begin) - t_phase2);
    return ILM_OK;
}

}  // namespace ilm

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


// =====================================================================================================================
// Gather lists on the device (the default; ILM_TABLES_HOST=1 keeps the host build below as the cross-check).
// One CTA per list: the four cell-bucketed gather lists (key = cell, value = k*W*W + slot) and the row buckets of the
// primal windows (key = row, value = k*W + b).  A CTA sorts its <= N*W*W pairs with a stable LSD radix sort, 8 bits per
// pass: thread t owns the contiguous chunk t of the sequence, counts its digits in a private column of a 256 x 128
// shared-memory histogram, the CTA scans the histogram in (digit, thread) order, and every thread scatters its chunk in
// order -- stable, no atomics, the same order as the host's radix sort (ties keep ascending point order = the order of a
// CSC column scan).  Run heads of the sorted keys give cell_idx / cell_off; row_ptr by binary search.
// =====================================================================================================================
constexpr int LB_THREADS = 128;
struct ListJob {
    int n, W, kind;                 // kind 0: cell list of a layout, 1: row buckets (primal windows)
    const int* i0; const int* j0;
    int mx, my;
    unsigned* key[2]; int* val[2];  // ping-pong buffers
    int* out_idx; int* out_off; int* out_ent;
    int* counts;                    // [0] = number of runs (cells), [1] = number of valid entries
};
struct ListJobs { ListJob j[5]; };

__device__ __forceinline__ int block_excl_scan_128(int v, int* s_w, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    int base = 0;
    for (int w2 = 0; w2 < wid; ++w2) base += s_w[w2];
    if (total) *total = s_w[0] + s_w[1] + s_w[2] + s_w[3];
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(LB_THREADS) k_build_lists(const __grid_constant__ ListJobs jobs) {
    extern __shared__ unsigned hist[];                       // [256][LB_THREADS]
    __shared__ int s_w[4];
    const ListJob& J = jobs.j[blockIdx.x];
    const int n = J.n, t = threadIdx.x;
    if (n <= 0) { if (t == 0) { J.counts[0] = 0; J.counts[1] = 0; if (J.kind == 0) J.out_off[0] = 0; } if (J.kind == 1) for (int j = t; j <= J.my; j += LB_THREADS) J.out_off[j] = 0; return; }
    const int m = (n + LB_THREADS - 1) / LB_THREADS;
    const int e0 = t * m, e1 = min(n, e0 + m);
    const unsigned invalid = J.kind == 0 ? (unsigned)J.mx * (unsigned)J.my : (unsigned)J.my;
    // keys and values in generation order (k, then b, then a)
    const int per = J.kind == 0 ? J.W * J.W : J.W;
    for (int e = e0; e < e1; ++e) {
        const int k = e / per, slot = e - k * per;
        unsigned key;
        if (J.kind == 0) {
            const int b = slot / J.W, a = slot - b * J.W;
            const int i = J.i0[k] + a, j = J.j0[k] + b;
            key = (i >= 0 && i < J.mx && j >= 0 && j < J.my) ? (unsigned)(i + J.mx * j) : invalid;
        } else {
            const int j = J.j0[k] + slot;
            key = (j >= 0 && j < J.my) ? (unsigned)j : invalid;
        }
        J.key[0][e] = key;
        J.val[0][e] = e;
    }
    int src = 0;
    for (unsigned shift = 0; shift < 32 && (invalid >> shift) > 0; shift += 8) {
        const unsigned* kin = J.key[src]; const int* vin = J.val[src];
        unsigned* kout = J.key[src ^ 1]; int* vout = J.val[src ^ 1];
        for (int d = 0; d < 256; ++d) hist[d * LB_THREADS + t] = 0;
        __syncthreads();                                     // the writes of the previous pass (and of the generation) are visible
        // the chunk is read in batches of 8 independent loads (a load per loop iteration, each chained to the shared-memory
        // increment, exposed 600 cycles of L2 latency per element: 2.1 ms per refresh)
        for (int e = e0; e < e1; e += 8) {
            unsigned kb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) kb[u] = e + u < e1 ? kin[e + u] : 0u;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (e + u < e1) hist[((kb[u] >> shift) & 255u) * LB_THREADS + t]++;
        }
        __syncthreads();
        // exclusive scan in (digit, thread) order: thread t scans digits 2t, 2t+1
        int tot = 0;
        for (int q = 0; q < 2 * LB_THREADS; ++q) tot += hist[(2 * t) * LB_THREADS + q];
        int base = block_excl_scan_128(tot, s_w, nullptr);
        for (int q = 0; q < 2 * LB_THREADS; ++q) { const unsigned c = hist[(2 * t) * LB_THREADS + q]; hist[(2 * t) * LB_THREADS + q] = base; base += c; }
        __syncthreads();
        for (int e = e0; e < e1; e += 8) {
            unsigned kb[8];
            int vb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { kb[u] = e + u < e1 ? kin[e + u] : 0u; vb[u] = e + u < e1 ? vin[e + u] : 0; }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (e + u < e1) {
                    const unsigned dst = hist[((kb[u] >> shift) & 255u) * LB_THREADS + t]++;
                    kout[dst] = kb[u];
                    vout[dst] = vb[u];
                }
        }
        src ^= 1;
        __syncthreads();
    }
    const unsigned* ks = J.key[src]; const int* vs = J.val[src];
    // valid entries come first; run heads
    int nv = 0, nh = 0;
    for (int e = e0; e < e1; ++e) {
        const unsigned key = ks[e];
        if (key != invalid) { ++nv; if (e == 0 || ks[e - 1] != key) ++nh; }
    }
    int nent = 0, ncell = 0;
    block_excl_scan_128(nv, s_w, &nent);
    int hbase = block_excl_scan_128(nh, s_w, &ncell);
    for (int e = e0; e < e1; ++e) {
        const unsigned key = ks[e];
        if (key != invalid) {
            J.out_ent[e] = vs[e];
            if (J.kind == 0 && (e == 0 || ks[e - 1] != key)) { J.out_idx[hbase] = (int)key; J.out_off[hbase] = e; ++hbase; }
        }
    }
    if (t == 0) { J.counts[0] = ncell; J.counts[1] = nent; if (J.kind == 0) J.out_off[ncell] = nent; }
    if (J.kind == 1) {
        // row_ptr[j] = first sorted position with row >= j
        for (int j = t; j <= J.my; j += LB_THREADS) {
            int lo = 0, hi = nent;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ks[mid] < (unsigned)j) lo = mid + 1; else hi = mid; }
            J.out_off[j] = lo;
        }
    }
}

// need[row*T + j] |= 1 << e for every window column i = j + e*T of every in-field window row (pass C column mask)
__global__ void k_need_mask(int N, int W, int mx, int my, int T, const int* __restrict__ i0, const int* __restrict__ j0, unsigned* __restrict__ need) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * W * W) return;
    const int k = e / (W * W), slot = e - k * W * W, b = slot / W, a = slot - b * W;
    const int i = i0[k] + a, j = j0[k] + b;
    if (i >= 0 && i < mx && j >= 0 && j < my) atomicOr(&need[(size_t)j * T + i % T], 1u << (i / T));
}

void free_table(DevTable& t) {
    cudaFree(t.i0); cudaFree(t.j0); cudaFree(t.wR); cudaFree(t.wE);
    cudaFree(t.cell_idx); cudaFree(t.cell_off); cudaFree(t.ent); cudaFree(t.rowsum);
    cudaFree(t.row_ptr); cudaFree(t.row_ent); cudaFree(t.part); cudaFree(t.need);
    for (int q = 0; q < 2; ++q) { cudaFree(t.skey[q]); cudaFree(t.sval[q]); cudaFree(t.rkey[q]); cudaFree(t.rval[q]); }
    cudaFree(t.dcounts);
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
    const bool host_lists = getenv("ILM_TABLES_HOST") != nullptr;       // read per call: the tests build both
    if (!host_lists) {
        // phase 2 on the device: all five lists by one launch, the column mask by another; one synchronisation for the
        // counts and the host mirror of j0
        ListJobs jobs{};
        const int T = p->Lx / 16;
        for (int layout = 0; layout < 4; ++layout) {
            DevTable& t = p->tab[layout];
            const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
            const size_t n = (size_t)N * W2;
            if (n + 16 > t.cap_sort || !t.skey[0]) {
                const size_t cap = n + n / 4 + 16;
                for (int q = 0; q < 2; ++q) {
                    cudaFree(t.skey[q]); cudaFree(t.sval[q]); t.skey[q] = nullptr; t.sval[q] = nullptr;
                    ILM_CUDA(cudaMalloc(&t.skey[q], cap * sizeof(unsigned)));
                    ILM_CUDA(cudaMalloc(&t.sval[q], cap * sizeof(int)));
                }
                cudaFree(t.cell_idx); cudaFree(t.cell_off); cudaFree(t.rowsum); cudaFree(t.ent);
                t.cell_idx = t.cell_off = t.ent = nullptr; t.rowsum = nullptr;
                ILM_CUDA(cudaMalloc(&t.cell_idx, cap * sizeof(int)));
                ILM_CUDA(cudaMalloc(&t.cell_off, (cap + 1) * sizeof(int)));
                ILM_CUDA(cudaMalloc(&t.rowsum, cap * sizeof(double)));
                ILM_CUDA(cudaMalloc(&t.ent, cap * sizeof(int)));
                t.cap_sort = cap; t.cap_cells = cap; t.cap_ents = cap;
            }
            if (!t.dcounts) ILM_CUDA(cudaMalloc(&t.dcounts, 4 * sizeof(int)));
            ListJob& J = jobs.j[layout];
            J.n = (int)n; J.W = W; J.kind = 0; J.i0 = t.i0; J.j0 = t.j0; J.mx = li.mx; J.my = li.my;
            J.key[0] = t.skey[0]; J.key[1] = t.skey[1]; J.val[0] = t.sval[0]; J.val[1] = t.sval[1];
            J.out_idx = t.cell_idx; J.out_off = t.cell_off; J.out_ent = t.ent; J.counts = t.dcounts;
        }
        {
            DevTable& t = p->tab[ILM_NODES_PRIMAL];
            const LayoutInfo li = layout_info(ILM_NODES_PRIMAL, p->g.NX, p->g.NY);
            const size_t n = (size_t)N * W;
            if ((size_t)li.my + 1 > t.cap_rows || !t.row_ptr) {
                cudaFree(t.row_ptr); t.row_ptr = nullptr;
                ILM_CUDA(cudaMalloc(&t.row_ptr, ((size_t)li.my + 1) * sizeof(int)));
                t.cap_rows = (size_t)li.my + 1;
            }
            if (np * W > t.cap_rowent || !t.row_ent || !t.rkey[0]) {
                cudaFree(t.row_ent); cudaFree(t.part); t.row_ent = nullptr; t.part = nullptr;
                const size_t cap = (np + np / 4 + 16) * W;
                ILM_CUDA(cudaMalloc(&t.row_ent, cap * sizeof(int)));
                ILM_CUDA(cudaMalloc(&t.part, cap * PART_BATCH * sizeof(double2)));
                for (int q = 0; q < 2; ++q) {
                    cudaFree(t.rkey[q]); cudaFree(t.rval[q]); t.rkey[q] = nullptr; t.rval[q] = nullptr;
                    ILM_CUDA(cudaMalloc(&t.rkey[q], cap * sizeof(unsigned)));
                    ILM_CUDA(cudaMalloc(&t.rval[q], cap * sizeof(int)));
                }
                t.part_stride = cap;
                t.cap_rowent = cap;
            }
            const size_t nneed = (size_t)li.my * T;
            if (nneed > t.cap_need || !t.need) {
                cudaFree(t.need); t.need = nullptr;
                ILM_CUDA(cudaMalloc(&t.need, (nneed + 4) * sizeof(unsigned)));
                t.cap_need = nneed;
            }
            t.need_T = T;
            ListJob& J = jobs.j[4];
            J.n = (int)n; J.W = W; J.kind = 1; J.i0 = t.i0; J.j0 = t.j0; J.mx = li.mx; J.my = li.my;
            J.key[0] = t.rkey[0]; J.key[1] = t.rkey[1]; J.val[0] = t.rval[0]; J.val[1] = t.rval[1];
            J.out_idx = nullptr; J.out_off = t.row_ptr; J.out_ent = t.row_ent; J.counts = t.dcounts + 2;
            ILM_CUDA(cudaMemsetAsync(t.need, 0, nneed * sizeof(unsigned), p->stream));
            if (N > 0) {
                k_need_mask<<<(N * W2 + 255) / 256, 256, 0, p->stream>>>(N, W, li.mx, li.my, T, t.i0, t.j0, t.need);
                p->launches++;
            }
        }
        static bool attr_done[64] = {};
        if (!attr_done[p->device & 63]) {
            ILM_CUDA(cudaFuncSetAttribute(k_build_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * LB_THREADS * (int)sizeof(unsigned)));
            attr_done[p->device & 63] = true;
        }
        k_build_lists<<<5, LB_THREADS, 256 * LB_THREADS * sizeof(unsigned), p->stream>>>(jobs);
        ILM_CUDA(cudaGetLastError());
        p->launches++;
        int hc[4][4];
        for (int layout = 0; layout < 4; ++layout)
            ILM_CUDA(cudaMemcpyAsync(hc[layout], p->tab[layout].dcounts, 4 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        ILM_CUDA(cudaStreamSynchronize(p->stream));
        for (int layout = 0; layout < 4; ++layout) {
            DevTable& t = p->tab[layout];
            t.h_j0 = hj[layout];
            t.ncell = hc[layout][0];
            t.nent = hc[layout][1];
        }
        ILM_TRY(launch_filter_rowsum(p, p->tab[ILM_NODES_PRIMAL]));
        if (ttrace) fprintf(stderr, "ilm table refresh, N = %d: %.3f ms (device lists)\n", N, ms_since(t_begin));
        return ILM_OK;
    }
    ILM_CUDA(cudaStreamSynchronize(p->stream));
    const double t_phase1 = ms_since(t_begin);
    // phase 2: gather lists (cell, k, slot) for every in-range window entry, sorted by cell then k.  Pure host work per
    // layout (entry generation, radix sort, run detection): the four layouts are built by four host threads, the uploads
    // follow in layout order.
    struct HostLists { std::vector<int> id, cell_idx, cell_off, rptr, rent; std::vector<unsigned> need; };
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
                        if (i >= 0 && i < li.mx) h.need[(size_t)j * T + i % T] |= 1u << (i / T);
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
                ILM_CUDA(cudaMalloc(&t.need, h.need.size() * sizeof(unsigned) + 16));
                t.cap_need = h.need.size();
            }
            t.need_T = p->Lx / 16;
            if (!h.need.empty()) ILM_CUDA(cudaMemcpyAsync(t.need, h.need.data(), h.need.size() * sizeof(unsigned), cudaMemcpyHostToDevice, p->stream));
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

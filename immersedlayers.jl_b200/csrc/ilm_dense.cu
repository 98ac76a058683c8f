// ilm_dense.cu -- dense surface-point solve: LU with partial pivoting, getrs,
// and the mat-vec power C^k s.  Replaces the LAPACK getrf/getrs behind `S\s`
// and `C^5*s` in the reference's user-level algorithm
// (test/literate/dirichlet.jl:99,124; neumann.jl:123,133).
//
// Right-looking blocked LU (NB = 32) with look-ahead on two prioritised streams.  The critical path is ONE kernel per
// panel (k_lu_panel_v2): a thread-block cluster holds the panel in registers (thread per row), applies the previous
// panel's update to its own 32 columns itself, and eliminates column by column with one cluster barrier per column,
// LAPACK's pivot choices, no row moves (logical row indices).  Beside it, on the low-priority stream, the wide update of
// the previous step: the panel's net row permutation as one gather + scatter per column with the block-row solve in
// between (k_lu_swap_trsm), and A22 -= A21 * A12 on the FP64 tensor pipe (k_lu_gemm: mma.sync.m8n8k4.f64, "DMMA").
// The earlier panel kernels (one CTA, shared-memory cluster, register panel with row moves) stay behind ILM_LU_V1 and
// for n > 8192; ILM_LU_TRACE prints the phase times of the panel kernel.  Triangular solves: k_getrs_cluster.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "ilm_internal.h"

namespace cg = cooperative_groups;

namespace ilm {

constexpr int NB = 32;
long long g_dense_launches = 0;

// ---- panel: columns [k0, k0+nb), rows [k0, n); one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_lu_panel(int n, int k0, int nb, double* __restrict__ A, int* __restrict__ ipiv) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double s_row[NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int jj = 0; jj < nb; ++jj) {
        const int col = k0 + jj;
        double* a_col = A + (size_t)col * n;
        // pivot search: first maximum of |a| (idamax semantics)
        double best = -1.0;
        int bi = n;
        for (int r = col + tid; r < n; r += 1024) {
            const double v = fabs(a_col[r]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            best = s_val[lane]; bi = s_idx[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_piv = bi; ipiv[col] = bi + 1; }
        }
        __syncthreads();
        const int piv = s_piv;
        // swap rows col <-> piv inside the panel, keep the pivot row in smem
        if (tid < nb) {
            double* c = A + (size_t)(k0 + tid) * n;
            const double top = c[piv];
            if (piv != col) { c[piv] = c[col]; c[col] = top; }
            s_row[tid] = top;
        }
        __syncthreads();
        const double pivot = s_row[jj];
        // scale the column and rank-1 update of the remaining panel columns
        const int ncols = nb - jj - 1;
        for (int r = col + 1 + tid; r < n; r += 1024) {
            const double l = a_col[r] / pivot;
            a_col[r] = l;
            for (int c = 0; c < ncols; ++c) {
                double* q = A + (size_t)(col + 1 + c) * n + r;
                *q = *q - l * s_row[jj + 1 + c];
            }
        }
        __syncthreads();
    }
}

// ---- panel on a thread-block cluster: 8 CTAs (8 SMs) share one panel, each owning a fixed slice
// of the rows below the diagonal block; the pivot candidates are exchanged through distributed
// shared memory, two cluster barriers per column.
constexpr int PANEL_CTAS = 8;
__global__ void __cluster_dims__(PANEL_CTAS, 1, 1) __launch_bounds__(1024)
k_lu_panel_cluster(int n, int k0, int nb, double* __restrict__ A, int* __restrict__ ipiv) {
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double c_val;          // this CTA's pivot candidate (read by the peers through DSMEM)
    __shared__ int c_idx;
    __shared__ int s_piv;
    __shared__ double s_row[NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // fixed row slices of [k0, n)
    const int len = n - k0, chunk = (len + PANEL_CTAS - 1) / PANEL_CTAS;
    const int sl0 = k0 + rank * chunk, sl1 = min(n, sl0 + chunk);
    for (int jj = 0; jj < nb; ++jj) {
        const int col = k0 + jj;
        double* a_col = A + (size_t)col * n;
        double best = -1.0;
        int bi = n;
        for (int r = max(sl0, col) + tid; r < sl1; r += 1024) {
            const double v = fabs(a_col[r]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            best = s_val[lane]; bi = s_idx[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { c_val = best; c_idx = bi; }
        }
        cl.sync();                                   // candidates of all CTAs are in place
        if (wid == 0) {
            best = -1.0; bi = n;
            if (lane < PANEL_CTAS) {
                best = *cl.map_shared_rank(&c_val, lane);
                bi = *cl.map_shared_rank(&c_idx, lane);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) s_piv = bi;
        }
        __syncthreads();
        const int piv = s_piv;
        if (rank == 0) {                             // row interchange inside the panel
            if (tid < nb) {
                double* c = A + (size_t)(k0 + tid) * n;
                const double top = c[piv];
                if (piv != col) { c[piv] = c[col]; c[col] = top; }
            }
            if (tid == 0) ipiv[col] = piv + 1;
            __threadfence();
        }
        cl.sync();                                   // the swapped pivot row is visible to every CTA
        if (tid < nb) s_row[tid] = A[(size_t)(k0 + tid) * n + col];
        __syncthreads();
        const double pivot = s_row[jj];
        const int ncols = nb - jj - 1;
        for (int r = max(sl0, col + 1) + tid; r < sl1; r += 1024) {
            const double l = a_col[r] / pivot;
            a_col[r] = l;
            for (int c = 0; c < ncols; ++c) {
                double* q = A + (size_t)(col + 1 + c) * n + r;
                *q = *q - l * s_row[jj + 1 + c];
            }
        }
        __syncthreads();
    }
    cl.sync();                                       // no CTA leaves while a peer may still read its smem
}

// ---- panel held in shared memory: each of the 8 CTAs of the cluster keeps its slice of the panel
// (rows x nb) in its shared memory for all nb column steps; pivot candidates, the pivot row and the
// row that is swapped out travel through distributed shared memory.  Per column: one block reduction,
// two cluster barriers, no global memory traffic.
__global__ void __cluster_dims__(PANEL_CTAS, 1, 1) __launch_bounds__(1024)
k_lu_panel_smem(int n, int k0, int nb, int ldp, double* __restrict__ A, int* __restrict__ ipiv) {
    extern __shared__ double P[];                    // [ldp][nb] column-major slice
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double c_val;
    __shared__ int c_idx;
    __shared__ int s_piv;
    __shared__ double s_row[NB], s_old[NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int len = n - k0, chunk = (len + PANEL_CTAS - 1) / PANEL_CTAS;
    const int sl0 = k0 + rank * chunk, sl1 = min(n, sl0 + chunk);
    const int nrow = max(sl1 - sl0, 0);
    for (int c = 0; c < nb; ++c)
        for (int r = tid; r < nrow; r += 1024) P[r + ldp * c] = A[(size_t)(k0 + c) * n + sl0 + r];
    __syncthreads();
    for (int jj = 0; jj < nb; ++jj) {
        const int col = k0 + jj;
        const double* pc = P + ldp * jj;
        double best = -1.0;
        int bi = n;
        for (int r = max(col - sl0, 0) + tid; r < nrow; r += 1024) {
            const double v = fabs(pc[r]);
            if (v > best) { best = v; bi = sl0 + r; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            best = s_val[lane]; bi = s_idx[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { c_val = best; c_idx = bi; }
        }
        cl.sync();                                   // (A) candidates in place, previous updates finished
        if (wid == 0) {
            best = -1.0; bi = n;
            if (lane < PANEL_CTAS) {
                best = *cl.map_shared_rank(&c_val, lane);
                bi = *cl.map_shared_rank(&c_idx, lane);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) s_piv = bi;
        }
        __syncthreads();
        const int piv = s_piv;
        const int rp = (piv - k0) / chunk, rc = (col - k0) / chunk;      // owners of the two rows
        if (tid < nb) {
            const double* Pp = cl.map_shared_rank(P, rp);
            s_row[tid] = Pp[(piv - (k0 + rp * chunk)) + ldp * tid];
        } else if (tid >= 32 && tid < 32 + nb) {
            const double* Pc = cl.map_shared_rank(P, rc);
            s_old[tid - 32] = Pc[(col - (k0 + rc * chunk)) + ldp * (tid - 32)];
        }
        cl.sync();                                   // (B) every CTA holds both rows; owners may overwrite
        if (piv != col && tid < nb) {
            if (rank == rc) P[(col - sl0) + ldp * tid] = s_row[tid];
            if (rank == rp) P[(piv - sl0) + ldp * tid] = s_old[tid];
        }
        if (rank == 0 && tid == 0) ipiv[col] = piv + 1;
        __syncthreads();
        const double pivot = s_row[jj];
        const int ncols = nb - jj - 1;
        double* pcw = P + ldp * jj;
        for (int r = max(col + 1 - sl0, 0) + tid; r < nrow; r += 1024) {
            const double l = pcw[r] / pivot;
            pcw[r] = l;
            for (int c = 0; c < ncols; ++c) {
                double* q = P + ldp * (jj + 1 + c) + r;
                *q = *q - l * s_row[jj + 1 + c];
            }
        }
        __syncthreads();
    }
    for (int c = 0; c < nb; ++c)
        for (int r = tid; r < nrow; r += 1024) A[(size_t)(k0 + c) * n + sl0 + r] = P[r + ldp * c];
    cl.sync();                                       // no CTA leaves while a peer may still read its smem
}

// ---- panel held in REGISTERS: one thread per row (512 threads x 8 or 16 CTAs of one cluster), the thread keeps
// its nb <= 32 panel entries in registers for all column steps.  Per column: a warp + CTA arg-max, the CTA's
// candidate row and the current top row are pushed into every peer's shared memory (DSMEM, double-buffered by
// column parity), ONE cluster barrier, then every thread picks the winner locally, swaps in registers and
// updates its row.  (The shared-memory version above needs two cluster barriers and four block barriers per column.)
constexpr int PR_THREADS = 512, PR_WARPS = PR_THREADS / 32, PR_MAXC = 16;
struct PanelXchg {
    double val[2][PR_MAXC];
    int idx[2][PR_MAXC];
    double rows[2][PR_MAXC][NB];
    double old[2][NB];
};
__device__ __forceinline__ void argmax_step(double& best, int& bi, int off) {
    const double ov = __shfl_down_sync(0xffffffffu, best, off);
    const int oi = __shfl_down_sync(0xffffffffu, bi, off);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
}
__global__ void __launch_bounds__(PR_THREADS) k_lu_panel_regs(int n, int k0, int nb, double* __restrict__ A, int* __restrict__ ipiv) {
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    __shared__ PanelXchg X;
    __shared__ double s_val[PR_WARPS];
    __shared__ int s_idx[PR_WARPS];
    __shared__ double c_row[NB], o_row[NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = k0 + rank * PR_THREADS + tid;
    const bool has = row < n;
    double a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = (has && c < nb) ? A[(size_t)(k0 + c) * n + row] : 0.0;
    cl.sync();                                       // every CTA of the cluster is running before the first DSMEM push
    // The column loop is fully unrolled so that the register array is addressed statically.  (A rolled loop with
    // predicated static indices was tried because ncu shows 58 % no_instructions stalls on the ~250 KB of straight-line
    // code: it spills 400 B per thread at the 128-register cap and is slower, N = 4593: 23.9 -> 31.5 ms.)
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) {
        if (jj < nb) {
            const int col = k0 + jj, buf = jj & 1;
            // candidate of this warp, of this CTA (every warp reduces the warp candidates redundantly)
            const bool active = has && row >= col;
            double best = active ? fabs(a[jj]) : -1.0;
            int bi = active ? row : n;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) argmax_step(best, bi, off);
            if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
            __syncthreads();
            best = lane < PR_WARPS ? s_val[lane] : -1.0;
            bi = lane < PR_WARPS ? s_idx[lane] : n;
#pragma unroll
            for (int off = PR_WARPS / 2; off > 0; off >>= 1) argmax_step(best, bi, off);
            const double cta_val = __shfl_sync(0xffffffffu, best, 0);
            const int cta_piv = __shfl_sync(0xffffffffu, bi, 0);
            if (has && row == cta_piv) {
#pragma unroll
                for (int c = 0; c < NB; ++c) c_row[c] = a[c];
            }
            if (has && row == col) {
#pragma unroll
                for (int c = 0; c < NB; ++c) o_row[c] = a[c];
            }
            __syncthreads();
            if (wid < C) {                               // warp p pushes to peer p
                PanelXchg* Xp = cl.map_shared_rank(&X, wid);
                Xp->rows[buf][rank][lane] = c_row[lane];
                if (lane == 0) { Xp->val[buf][rank] = cta_val; Xp->idx[buf][rank] = cta_piv; }
                if (rank == (col - k0) / PR_THREADS) Xp->old[buf][lane] = o_row[lane];
            }
            cl.sync();
            double bv = -1.0;
            int piv = n, pw = 0;
            for (int p = 0; p < C; ++p) {
                const double v = X.val[buf][p];
                const int i = X.idx[buf][p];
                if (v > bv || (v == bv && i < piv)) { bv = v; piv = i; pw = p; }
            }
            const double* prow = X.rows[buf][pw];
            if (piv != col) {
                if (has && row == piv) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) a[c] = X.old[buf][c];
                } else if (has && row == col) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) a[c] = prow[c];
                }
            }
            if (rank == 0 && tid == 0) ipiv[col] = piv + 1;
            if (has && row > col) {
                const double l = a[jj] / prow[jj];
                a[jj] = l;
#pragma unroll
                for (int c = jj + 1; c < NB; ++c) a[c] = a[c] - l * prow[c];
            }
        }
    }
    if (has) {
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) A[(size_t)(k0 + c) * n + row] = a[c];
    }
    cl.sync();                                       // no CTA leaves while a peer may still write its shared memory
}

__global__ void __launch_bounds__(PR_THREADS) k_lu_panel_rot(int n, int k0, int nb, double* __restrict__ A, int* __restrict__ ipiv) {
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    __shared__ PanelXchg X;
    __shared__ double s_val[PR_WARPS];
    __shared__ int s_idx[PR_WARPS];
    __shared__ double c_row[NB], o_row[NB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = k0 + rank * PR_THREADS + tid;
    const bool has = row < n;
    double a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = (has && c < nb) ? A[(size_t)(k0 + c) * n + row] : 0.0;
    cl.sync();                                       // every CTA of the cluster is running before the first DSMEM push
    // ROLLED column loop: the register row is rotated by one position per column, so the current column is always
    // a[0] and the loop body is the same code for every column (the fully unrolled form, k_lu_panel_regs, is ~250 KB of
    // straight-line code: ncu showed 58 % no_instructions stalls).  Position p holds logical column (jj + p) mod 32;
    // positions >= 32 - jj are the finished multipliers of the columns to the left.  All threads rotate alike, so the
    // row exchanges through shared memory stay position-wise.
#pragma unroll 1
    for (int jj = 0; jj < NB; ++jj) {
        if (jj < nb) {
            const int col = k0 + jj, buf = jj & 1;
            // candidate of this warp, of this CTA (every warp reduces the warp candidates redundantly)
            const bool active = has && row >= col;
            double best = active ? fabs(a[0]) : -1.0;
            int bi = active ? row : n;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) argmax_step(best, bi, off);
            if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
            __syncthreads();
            best = lane < PR_WARPS ? s_val[lane] : -1.0;
            bi = lane < PR_WARPS ? s_idx[lane] : n;
#pragma unroll
            for (int off = PR_WARPS / 2; off > 0; off >>= 1) argmax_step(best, bi, off);
            const double cta_val = __shfl_sync(0xffffffffu, best, 0);
            const int cta_piv = __shfl_sync(0xffffffffu, bi, 0);
            if (has && row == cta_piv) {
#pragma unroll
                for (int c = 0; c < NB; ++c) c_row[c] = a[c];
            }
            if (has && row == col) {
#pragma unroll
                for (int c = 0; c < NB; ++c) o_row[c] = a[c];
            }
            __syncthreads();
            if (wid < C) {                               // warp p pushes to peer p
                PanelXchg* Xp = cl.map_shared_rank(&X, wid);
                Xp->rows[buf][rank][lane] = c_row[lane];
                if (lane == 0) { Xp->val[buf][rank] = cta_val; Xp->idx[buf][rank] = cta_piv; }
                if (rank == (col - k0) / PR_THREADS) Xp->old[buf][lane] = o_row[lane];
            }
            cl.sync();
            double bv = -1.0;
            int piv = n, pw = 0;
            for (int p = 0; p < C; ++p) {
                const double v = X.val[buf][p];
                const int i = X.idx[buf][p];
                if (v > bv || (v == bv && i < piv)) { bv = v; piv = i; pw = p; }
            }
            const double* prow = X.rows[buf][pw];
            if (piv != col) {
                if (has && row == piv) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) a[c] = X.old[buf][c];
                } else if (has && row == col) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) a[c] = prow[c];
                }
            }
            if (rank == 0 && tid == 0) ipiv[col] = piv + 1;
            if (has && row > col) {
                const double l = a[0] / prow[0];
                a[0] = l;
                const int live = NB - jj;                    // positions 1 .. live-1 are the columns still to eliminate
#pragma unroll
                for (int c = 1; c < NB; ++c)
                    if (c < live) a[c] = a[c] - l * prow[c];
            }
            {                                                // rotate: the finished column goes to the back
                const double t0 = a[0];
#pragma unroll
                for (int c = 0; c < NB - 1; ++c) a[c] = a[c + 1];
                a[NB - 1] = t0;
            }
        }
    }
    if (has) {                                           // after nb rotations position p holds logical column (p + nb) mod 32
        const int done = nb < NB ? nb : NB;
#pragma unroll
        for (int p = 0; p < NB; ++p) {
            const int c = (p + done) & (NB - 1);
            if (c < nb) A[(size_t)(k0 + c) * n + row] = a[p];
        }
    }
    cl.sync();                                       // no CTA leaves while a peer may still write its shared memory
}

static int launch_panel_regs(int n, int k0, int nb, double* A, int* ipiv, cudaStream_t st) {
    const int rows = n - k0;
    int C = rows <= 8 * PR_THREADS ? 8 : 16;
    static const int adapt = getenv("ILM_LU_PANEL_ADAPT") ? atoi(getenv("ILM_LU_PANEL_ADAPT")) : 0;   // smallest cluster that holds the rows
    if (adapt > 0) { C = adapt; while (C * PR_THREADS < rows) C *= 2; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(PR_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static const bool unrolled = getenv("ILM_LU_PANEL_UNROLLED") != nullptr;     // the fully unrolled panel, kept for comparison
    if (unrolled) ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_regs, n, k0, nb, A, ipiv));
    else ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_rot, n, k0, nb, A, ipiv));
    return ILM_OK;
}

// =====================================================================================================================
// Panel, second form (the default): the same register panel (thread per row, cluster of C CTAs x 512 threads) with
//   * the NARROW UPDATE of the previous panel fused in front: the kernel itself applies panel k-1's interchanges to its
//     32 columns (each thread loads the row its own row ends up holding), one warp solves the 32 x 32 block row
//     U12 = L11^-1 A12, the block is pushed to every CTA of the cluster and every thread subtracts L21 U12 from its
//     row in registers.  Replaces three dependent launches (swap, trsm, rank-32 DMMA update of 32 columns: ~35 us)
//     on the critical path of the factorisation by ~8 us inside the panel kernel;
//   * arg-max by three redux.sync on the ordered bit pattern of |a| (max of the high word, max of the low word among
//     those, min row index among those = first maximum, idamax semantics) instead of five shuffle rounds of
//     (value, index) pairs;
//   * every warp's candidate row written to shared memory BEFORE the block barrier (speculatively), so one block
//     barrier per column instead of two; the final 16-candidate scan is lane-parallel (redux again);
//   * the reciprocal of the candidate pivot computed by the pushing warp while the row travels (LAPACK's getf2
//     scales by the reciprocal too).
// Pivot choices are those of LAPACK (first maximum of |a|); the multipliers l = a * (1/p).
// =====================================================================================================================
constexpr int P2_MAXC = 16;
struct P2Xchg {                        // double-buffered by column parity
    double rows[2][P2_MAXC][NB];       // candidate row of every CTA
    double ext[2][P2_MAXC][4];         // [0] 1 / candidate, [1] key (|a| bit pattern), [2] logical row index (as bits)
};
constexpr int P2_NOROW = 0x7fffffff;
// Maximum of the 64-bit keys (hi, lo) over the lanes with `valid`, ties to the smallest idx (idamax semantics on the logical
// row order).  One redux.sync + one ballot unless several lanes share the high word.  Returns the winning lane or -1.
__device__ __forceinline__ int first_max_lane(unsigned hi, unsigned lo, int idx, bool valid) {
    const unsigned mh = __reduce_max_sync(0xffffffffu, valid ? hi : 0u);
    unsigned m = __ballot_sync(0xffffffffu, valid && hi == mh);
    if (__popc(m) > 1) {
        const bool in = valid && hi == mh;
        const unsigned ml = __reduce_max_sync(0xffffffffu, in ? lo : 0u);
        m = __ballot_sync(0xffffffffu, in && lo == ml);
        if (__popc(m) > 1) {                                 // equal |a|: the smallest (logical) row index wins
            const bool in2 = in && lo == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, in2 ? (unsigned)idx : 0xffffffffu);
            m = __ballot_sync(0xffffffffu, in2 && (unsigned)idx == mi);
        }
    }
    return __ffs(m) - 1;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// per panel: the net row permutation of its interchanges, for the wide update (k_lu_swap_trsm).
// perm[0] = m, perm[1 + q] = dst row, perm[1 + 64 + q] = src row; the first nb entries are the top block rows in order.
constexpr int PERM_STRIDE = 132;

#define P2A(p) (((p) & 1) ? a[(p) >> 1].y : a[(p) >> 1].x)

template <int THREADS, int GS>
__global__ void __launch_bounds__(THREADS) k_lu_panel_v2(int n, int k0, int nb, double* __restrict__ A, int* __restrict__ ipiv,
                                                            int pk0, int pnb, int* __restrict__ perm, double* __restrict__ uscr,
                                                            long long* __restrict__ trace) {
    // trace (ILM_LU_TRACE=1, null otherwise): clock64 sums of the phases of the column loop seen by thread 0 of CTA 0
    __shared__ long long s_tr[16];
    long long tlast = 0;
    auto stamp = [&](int slot) {
        if (trace && threadIdx.x == 0 && blockIdx.x == 0) {
            const long long t = clock64();
            if (slot >= 0) s_tr[slot] += t - tlast; else { for (int i = 0; i < 16; ++i) s_tr[i] = 0; }
            tlast = clock64();
        }
    };
    stamp(-1);
    if (trace && threadIdx.x == 0 && blockIdx.x == 0) {          // timeline: entry / exit of every panel kernel (ns)
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace[16 + 2 * (k0 / NB)] = (long long)t;
    }
    constexpr int WARPS = THREADS / 32;
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    __shared__ __align__(16) P2Xchg X;
    __shared__ __align__(16) double s_rows[WARPS][NB];
    __shared__ unsigned s_khi[WARPS], s_klo[WARPS];
    __shared__ int s_idx[WARPS];
    __shared__ int s_piv[NB];
    __shared__ double s_ext[4];
    __shared__ __align__(16) double Ubuf[NB][NB];            // U12 of the fused narrow update, [j][c]
    __shared__ double Tt[NB][NB + 1], LL[NB][NB + 1];        // rank 0, warp 0: transposition tile and L11
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = k0 + rank * THREADS + tid;
    const bool has = row < n;
    double2 a[NB / 2];                                       // the row: position p = a[p/2].x / .y (vector shared-memory accesses)
    if (pnb > 0) {
        // ---- fused narrow update with panel [pk0, pk0 + pnb), pk0 + pnb == k0
        if (tid < NB) s_piv[tid] = tid < pnb ? ipiv[pk0 + tid] - 1 : -1;
        __syncthreads();
        // the row that ends up in row r after the interchanges jj = 0 .. pnb-1 (applied backwards to the index)
        auto source_row = [&](int r) {
            for (int jj = pnb - 1; jj >= 0; --jj) {
                const int t = pk0 + jj, pv = s_piv[jj];
                r = (r == t) ? pv : ((r == pv) ? t : r);
            }
            return r;
        };
        // rows >= k0 move only if they were a pivot of the previous panel (independent compares; the dependent
        // 32-step walk only for those)
        bool moved = false;
#pragma unroll 8
        for (int jj = 0; jj < NB; ++jj) moved |= (row == s_piv[jj]);
        const int src = (has && moved) ? source_row(row) : (has ? row : 0);
        if (rank == 0 && wid == 0) {
            // block row: lane i holds row pk0 + i of the 32 columns; transpose through shared memory so that lane c
            // solves column c by forward substitution with the unit lower triangle L11
            const int ur = pk0 + lane;
            const int usrc = lane < pnb ? source_row(ur) : -1;
#pragma unroll
            for (int c = 0; c < NB; ++c) Tt[lane][c] = (usrc >= 0 && c < nb) ? A[(size_t)(k0 + c) * n + usrc] : 0.0;
#pragma unroll
            for (int j = 0; j < NB; ++j) LL[lane][j] = (lane < pnb && j < lane) ? A[(size_t)(pk0 + j) * n + ur] : 0.0;
            __syncwarp();
            double x[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) x[i] = Tt[i][lane];
            // column-oriented substitution: for fixed j the updates of x[j+1..] are independent (the row-oriented order
            // is one dependent chain of 496 FMAs)
#pragma unroll
            for (int j = 0; j < NB - 1; ++j) {
#pragma unroll
                for (int i = j + 1; i < NB; ++i) x[i] -= LL[i][j] * x[j];
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) Ubuf[i][lane] = x[i];
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) P2A(c) = (has && c < nb) ? A[(size_t)(k0 + c) * n + src] : 0.0;
        double l[2][8];                                          // multipliers of the previous panel, 8 at a time, one chunk ahead
#pragma unroll
        for (int t = 0; t < 8; ++t) l[0][t] = (has && t < pnb) ? A[(size_t)(pk0 + t) * n + row] : 0.0;
        stamp(11);
        __syncthreads();
        stamp(12);
        // U12 reaches the other CTAs through L2 (a scratch block, not the matrix: rows of the matrix may still be read as
        // interchange sources): pushing 8 KB to 15 peers through distributed shared memory cost 4.8 k cycles, bound by the
        // egress of one SM
        if (rank == 0 && C > 1) {
            const double2* s2 = reinterpret_cast<const double2*>(&Ubuf[0][0]);
            double2* g2 = reinterpret_cast<double2*>(uscr);
            for (int e = tid; e < NB * NB / 2; e += THREADS) g2[e] = s2[e];
        }
        cl.sync();                                               // U12 in L2; every load of the old rows has completed
        if (rank != 0) {
            double2* d2 = reinterpret_cast<double2*>(&Ubuf[0][0]);
            const double2* g2 = reinterpret_cast<const double2*>(uscr);
            for (int e = tid; e < NB * NB / 2; e += THREADS) d2[e] = __ldcg(g2 + e);
            __syncthreads();
        }
        stamp(13);
        if (rank == 0) {                                         // the block row goes back to the matrix
            for (int e = tid; e < NB * NB; e += THREADS) {
                const int j = e & (NB - 1), c = e >> 5;
                if (j < pnb && c < nb) A[(size_t)(k0 + c) * n + pk0 + j] = Ubuf[j][c];
            }
        }
#pragma unroll
        for (int ch = 0; ch < NB / 8; ++ch) {
            if (ch + 1 < NB / 8) {
#pragma unroll
                for (int t = 0; t < 8; ++t) l[(ch + 1) & 1][t] = (has && 8 * (ch + 1) + t < pnb) ? A[(size_t)(pk0 + 8 * (ch + 1) + t) * n + row] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double2* u = reinterpret_cast<const double2*>(&Ubuf[8 * ch + t][0]);
                const double lt = l[ch & 1][t];
#pragma unroll
                for (int c2 = 0; c2 < NB / 2; ++c2) {
                    const double2 uu = u[c2];
                    a[c2].x -= lt * uu.x;
                    a[c2].y -= lt * uu.y;
                }
            }
        }
        stamp(14);
    } else {
#pragma unroll
        for (int c = 0; c < NB; ++c) P2A(c) = (has && c < nb) ? A[(size_t)(k0 + c) * n + row] : 0.0;
        cl.sync();                                               // every CTA of the cluster is running before the first push
    }
    stamp(0);
    // ---- the columns, in 4 groups of 8: inside a group the sub-step s is a compile-time constant (current column at
    // position s, static register indices, no moves), after each group the register row is rotated by 8 positions, so the
    // loop body is the same code for every group and after the 4 rotations position p holds column p again.  In group g
    // the positions [32 - 8g, 32) hold finished multipliers of earlier groups: the 8-position blocks b >= 4 - g are skipped.
    //
    // NO ROW MOVES inside the panel: a thread keeps its row for the whole kernel and tracks the LOGICAL row index `pos`
    // that LAPACK's interchanges would have given it (the top row that is not chosen takes the pivot's index, the pivot
    // takes the column index and is done: neither searched nor updated any more).  Ties of |a| go to the smallest
    // logical index (idamax on the interchanged matrix).  At the end every thread writes its row to row `pos`.
    bool done = !has;
    int pos = row;
    int buf = 0;
#pragma unroll 1
    for (int g = 0; g < NB / GS; ++g) {
        const int liveb = NB / GS - g;                           // live GS-position blocks
        auto substep = [&](auto sconst) {
            constexpr int sidx = decltype(sconst)::value;
            const int jj = GS * g + sidx;
            if (jj < nb) {
                const int col = k0 + jj;
                const double av = fabs(P2A(sidx));
                const unsigned hi = (unsigned)__double2hiint(av), lo = (unsigned)__double2loint(av);
                const int w = first_max_lane(hi, lo, pos, !done);
                stamp(1);
                if (lane == w) { s_khi[wid] = hi; s_klo[wid] = lo; s_idx[wid] = pos; }
                if (w < 0 && lane == 0) { s_khi[wid] = 0u; s_klo[wid] = 0u; s_idx[wid] = P2_NOROW; }
                stamp(2);
                __syncthreads();
                stamp(3);
                // every warp finds the CTA's candidate; its owner alone writes the live part of its row to shared memory
                const bool l16 = lane < WARPS;
                const int myi = l16 ? s_idx[lane] : P2_NOROW;
                const unsigned mhi = l16 ? s_khi[lane] : 0u, mlo = l16 ? s_klo[lane] : 0u;
                const int ww = first_max_lane(mhi, mlo, myi, myi != P2_NOROW);
                if (wid == ww && lane == w) {
                    double2* d = reinterpret_cast<double2*>(&s_rows[0][0]);
#pragma unroll
                    for (int c2 = 0; c2 < NB / 2; ++c2)
                        if (2 * c2 + 1 > sidx && (2 * c2 / GS) < liveb) d[c2] = a[c2];
                    s_ext[0] = __drcp_rn(P2A(sidx));
                    s_ext[1] = av;
                    s_ext[2] = __longlong_as_double((long long)pos);
                }
                if (ww < 0 && tid == 0) { s_ext[1] = 0.0; s_ext[2] = __longlong_as_double((long long)P2_NOROW); }   // no row left in this CTA
                stamp(4);
                __syncthreads();
                {
                    const double rv = s_rows[0][lane], ev = s_ext[lane & 3];
                    for (int peer = wid; peer < C; peer += WARPS) {          // warp p pushes the CTA's candidate to the peers p, p + WARPS, ...
                        P2Xchg* Xp = cl.map_shared_rank(&X, peer);
                        Xp->rows[buf][rank][lane] = rv;
                        if (lane < 3) Xp->ext[buf][rank][lane] = ev;
                    }
                }
                stamp(5);
                __syncwarp();
                cluster_arrive();
                cluster_wait();
                stamp(6);
                const bool lc = lane < C;
                const double ck = lc ? X.ext[buf][lane][1] : 0.0;
                const int cli = lc ? (int)__double_as_longlong(X.ext[buf][lane][2]) : P2_NOROW;
                const int pw = first_max_lane((unsigned)__double2hiint(ck), (unsigned)__double2loint(ck), cli, cli != P2_NOROW);
                const int pivpos = (int)__double_as_longlong(X.ext[buf][pw][2]);     // LAPACK's pivot index of this column
                const double* prow = X.rows[buf][pw];
                stamp(7);
                if (rank == 0 && tid == 0) { ipiv[col] = pivpos + 1; s_piv[jj] = pivpos; }
                if (!done) {
                    if (pos == pivpos) { done = true; pos = col; }           // the pivot row: U row `col`, frozen
                    else if (pos == col) pos = pivpos;                       // the displaced top row
                }
                if (!done) {
                    const double lm = P2A(sidx) * X.ext[buf][pw][0];
                    P2A(sidx) = lm;
                    const double2* pr = reinterpret_cast<const double2*>(prow);
#pragma unroll
                    for (int c2 = 0; c2 < NB / 2; ++c2) {
                        if (2 * c2 + 1 > sidx && (2 * c2 / GS) < liveb) {
                            const double2 pp = pr[c2];
                            if (2 * c2 > sidx) a[c2].x -= lm * pp.x;
                            a[c2].y -= lm * pp.y;
                        }
                    }
                }
                buf ^= 1;
                stamp(8);
            }
        };
        substep(std::integral_constant<int, 0>{}); substep(std::integral_constant<int, 1>{});
        substep(std::integral_constant<int, 2>{}); substep(std::integral_constant<int, 3>{});
        if constexpr (GS == 8) {
            substep(std::integral_constant<int, 4>{}); substep(std::integral_constant<int, 5>{});
            substep(std::integral_constant<int, 6>{}); substep(std::integral_constant<int, 7>{});
        }
        {
            double2 t[GS / 2];                                   // rotate by GS: this group's multipliers go to the back
#pragma unroll
            for (int c = 0; c < GS / 2; ++c) t[c] = a[c];
#pragma unroll
            for (int c = 0; c < NB / 2 - GS / 2; ++c) a[c] = a[c + GS / 2];
#pragma unroll
            for (int c = 0; c < GS / 2; ++c) a[NB / 2 - GS / 2 + c] = t[c];
        }
    }
    stamp(9);
    if (has) {                                                   // four rotations by 8: position p holds column p again
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < nb) A[(size_t)(k0 + c) * n + pos] = P2A(c);
    }
    // ---- net permutation of this panel's interchanges for the wide update: rows touched = the top block rows and the
    // distinct pivot rows below it; each takes the value of the row its index maps back to
    __syncthreads();                                             // s_piv (thread 0) is complete
    if (rank == 0 && wid == 1 && perm) {
        const int pv = lane < nb ? s_piv[lane] : -1;
        const bool below = pv >= k0 + nb;
        const unsigned same = __match_any_sync(0xffffffffu, below ? pv : -1 - lane);
        const bool first = below && (same & ((1u << lane) - 1u)) == 0u;
        const unsigned fb = __ballot_sync(0xffffffffu, first);
        const int m = nb + __popc(fb);
        auto source_of = [&](int r) {
            for (int jj = nb - 1; jj >= 0; --jj) {
                const int t = k0 + jj, q = s_piv[jj];
                r = (r == t) ? q : ((r == q) ? t : r);
            }
            return r;
        };
        if (lane == 0) perm[0] = m;
        if (lane < nb) { perm[1 + lane] = k0 + lane; perm[1 + 64 + lane] = source_of(k0 + lane); }
        if (first) {
            const int q = nb + __popc(fb & ((1u << lane) - 1u));
            perm[1 + q] = pv; perm[1 + 64 + q] = source_of(pv);
        }
    }
    stamp(10);
    if (trace && threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i < 16; ++i) trace[i] += s_tr[i];
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace[17 + 2 * (k0 / NB)] = (long long)t;
    }
    cl.sync();                                                   // no CTA leaves while a peer may still write its shared memory
}
#undef P2A

static int launch_panel_v2(int n, int k0, int nb, double* A, int* ipiv, int pk0, int pnb, int* perm, double* uscr, long long* trace, cudaStream_t st) {
    const int rows = n - k0;
    // 256 threads per CTA whenever 16 CTAs hold the rows: twice as many SMs share the per-row work (the rank-32 update
    // of the fused narrow update is bound by the FP64 pipes of the cluster's SMs) and the block barriers see 8 warps
    static const int tpb_env = getenv("ILM_LU_PANEL_TPB") ? atoi(getenv("ILM_LU_PANEL_TPB")) : 0;
    const int tpb = tpb_env == 512 ? 512 : (rows <= P2_MAXC * 256 ? 256 : 512);
    int C = 1;
    while (C * tpb < rows) C *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(tpb);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    // GS = columns per unrolled group of the column loop (ILM_LU_PANEL_GS): 8 halves the register rotations, 4 halves the code
    // (ncu: 36 % no_instructions stalls on the 8-column body; measured 10.86 vs 11.08 ms at N = 4593)
    static const int gs = getenv("ILM_LU_PANEL_GS") ? atoi(getenv("ILM_LU_PANEL_GS")) : 4;
    if (tpb == 256) {
        if (gs == 4) ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_v2<256, 4>, n, k0, nb, A, ipiv, pk0, pnb, perm, uscr, trace));
        else ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_v2<256, 8>, n, k0, nb, A, ipiv, pk0, pnb, perm, uscr, trace));
    } else {
        if (gs == 4) ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_v2<512, 4>, n, k0, nb, A, ipiv, pk0, pnb, perm, uscr, trace));
        else ILM_CUDA(cudaLaunchKernelEx(&cfg, k_lu_panel_v2<512, 8>, n, k0, nb, A, ipiv, pk0, pnb, perm, uscr, trace));
    }
    return ILM_OK;
}

// Wide update, first kernel: the interchanges of panel [k0, k0 + nb) applied to the columns [0, k0) and [cright, n) as ONE
// gather and ONE scatter per column (the net permutation comes from the panel kernel; 32 dependent swaps cost 32 global
// round trips), and for the right columns the block row U12 = L11^-1 A12 solved in between, in registers.
__global__ void __launch_bounds__(64) k_lu_swap_trsm(int n, int k0, int nb, double* __restrict__ A, const int* __restrict__ perm,
                                                     int cright) {
    __shared__ int s_dst[64], s_src[64];
    __shared__ int s_m;
    __shared__ double L11[NB][NB + 1];
    const int nleft = k0;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) s_m = perm[0];
    if (threadIdx.x < 64) { s_dst[threadIdx.x] = perm[1 + threadIdx.x]; s_src[threadIdx.x] = perm[1 + 64 + threadIdx.x]; }
    const bool right_block = (blockIdx.x + 1) * blockDim.x > nleft;     // some column of this block needs the solve
    if (right_block)
        for (int i = threadIdx.x; i < NB * NB; i += blockDim.x) {
            const int r = i % NB, q = i / NB;
            L11[r][q] = (r < nb && q < r) ? A[(size_t)(k0 + q) * n + k0 + r] : 0.0;
        }
    __syncthreads();
    const bool right = c >= nleft;
    if (right) c += cright - nleft;
    if (c >= n) return;
    const int m = s_m;
    double* col = A + (size_t)c * n;
    double x[NB], y[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? col[s_src[i]] : 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) y[q] = nb + q < m ? col[s_src[nb + q]] : 0.0;
    if (right) {                                        // column-oriented substitution (independent updates for fixed j)
#pragma unroll
        for (int j = 0; j < NB - 1; ++j) {
#pragma unroll
            for (int i = j + 1; i < NB; ++i) x[i] -= L11[i][j] * x[j];
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (i < nb) col[k0 + i] = x[i];
#pragma unroll
    for (int q = 0; q < NB; ++q)
        if (nb + q < m) col[s_dst[nb + q]] = y[q];
}

// apply the row interchanges of pivots [r0, r0+nr) to the columns [cbeg, cend) except [skip0, skip1)
__global__ void k_lu_swap(int n, int r0, int nr, double* __restrict__ A, const int* __restrict__ ipiv, int cbeg, int cend,
                          int skip0, int skip1) {
    int c = cbeg + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= skip0) c += skip1 - skip0;
    if (c >= cend) return;
    double* col = A + (size_t)c * n;
    for (int jj = 0; jj < nr; ++jj) {
        const int r = r0 + jj, pv = ipiv[r] - 1;
        if (pv != r) { const double t = col[r]; col[r] = col[pv]; col[pv] = t; }
    }
}
static int launch_swap(int n, int r0, int nr, double* A, const int* ipiv, int cbeg, int cend, int skip0, int skip1, cudaStream_t st) {
    if (skip0 < cbeg) skip0 = cbeg;
    if (skip1 > cend) skip1 = cend;
    if (skip1 < skip0) skip1 = skip0;
    const int ncol = (cend - cbeg) - (skip1 - skip0);
    if (ncol <= 0 || nr <= 0) return ILM_OK;
    // (a gather / replay-in-shared-memory / scatter form of the interchanges was measured slower: N = 4593 LU 24.1 -> 26.4 ms)
    k_lu_swap<<<(ncol + 127) / 128, 128, 0, st>>>(n, r0, nr, A, ipiv, cbeg, cend, skip0, skip1);
    g_dense_launches++;
    return ILM_OK;
}

// block row: A12 <- L11^-1 A12 (unit lower triangular nb x nb)
__global__ void k_lu_trsm(int n, int k0, int nb, double* __restrict__ A, int cbeg, int cend) {
    __shared__ double L11[NB][NB + 1];
    for (int i = threadIdx.x; i < nb * nb; i += blockDim.x) L11[i % nb][i / nb] = A[(size_t)(k0 + i / nb) * n + k0 + i % nb];
    __syncthreads();
    const int c = cbeg + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cend) return;
    double* col = A + (size_t)c * n + k0;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? col[i] : 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) x[i] -= L11[i][j] * x[j];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (i < nb) col[i] = x[i];
}

// trailing update C -= A * B, K = kk; 64x64 tile per CTA, 4 warps of 32x32, DMMA m8n8k4
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) k_lu_gemm(int M, int Nc, int kk, const double* __restrict__ A,
                                                 const double* __restrict__ B, double* __restrict__ C, int ld) {
    constexpr int LDA = 68, LDB = 36;
    __shared__ double As[NB * LDA];      // As[k][m]
    __shared__ double Bs[64 * LDB];      // Bs[n][k]
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kc = 0; kc < kk; kc += NB) {            // K in chunks of 32 (one chunk for the inner updates)
        if (kc) __syncthreads();
        for (int i = tid; i < 64 * NB; i += 128) {
            const int m = i & 63, k = i >> 6;
            As[k * LDA + m] = (m0 + m < M && kc + k < kk) ? A[(size_t)(kc + k) * ld + m0 + m] : 0.0;
        }
        for (int i = tid; i < 64 * NB; i += 128) {
            const int k = i & (NB - 1), nn = i >> 5;
            Bs[nn * LDB + k] = (n0 + nn < Nc && kc + k < kk) ? B[(size_t)(n0 + nn) * ld + kc + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NB; k += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(k + t) * LDA + wm + 8 * i + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wn + 8 * j + g) * LDB + k + t];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    // all 32 loads of C are issued before the first store: a load-subtract-store per element would serialise 32
    // global round trips (the compiler cannot move a load of C above an earlier store to C)
    double cv[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + wm + 8 * i + g;
            const int c = n0 + wn + 8 * j + 2 * t;
            cv[i][j][0] = (r < M && c < Nc) ? C[(size_t)c * ld + r] : 0.0;
            cv[i][j][1] = (r < M && c + 1 < Nc) ? C[(size_t)(c + 1) * ld + r] : 0.0;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + wm + 8 * i + g;
            const int c = n0 + wn + 8 * j + 2 * t;
            if (r < M) {
                if (c < Nc) C[(size_t)c * ld + r] = cv[i][j][0] - acc[i][j][0];
                if (c + 1 < Nc) C[(size_t)(c + 1) * ld + r] = cv[i][j][1] - acc[i][j][1];
            }
        }
}

struct DenseIo {      // host/device staging without a plan
    cudaStream_t st;
    std::vector<void*> owned;
    struct Back { void* host; const void* dev; size_t bytes; };
    std::vector<Back> back;
    explicit DenseIo(void* stream) : st((cudaStream_t)stream) {}
    ~DenseIo() { for (void* q : owned) cudaFree(q); }
    template <class T> int map(T* user, size_t n, bool copy_in, bool copy_out, T** dev) {
        if (is_device_ptr(user)) { *dev = user; return ILM_OK; }
        void* d = nullptr;
        ILM_CUDA(cudaMalloc(&d, (n ? n : 1) * sizeof(T)));
        owned.push_back(d);
        if (copy_in) ILM_CUDA(cudaMemcpyAsync(d, user, n * sizeof(T), cudaMemcpyHostToDevice, st));
        if (copy_out) back.push_back({(void*)user, d, n * sizeof(T)});
        *dev = (T*)d;
        return ILM_OK;
    }
    int finish() {
        for (auto& b : back) ILM_CUDA(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st));
        if (!back.empty()) ILM_CUDA(cudaStreamSynchronize(st));
        return ILM_OK;
    }
};

// ---- getrs pieces: single CTA, right-hand side in shared memory
__global__ void __launch_bounds__(1024) k_getrs(int n, const double* __restrict__ LU, const int* __restrict__ ipiv,
                                                double* __restrict__ b) {
    extern __shared__ double sb[];
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 1024) sb[i] = b[i];
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < n; ++i) {
            const int pv = ipiv[i] - 1;
            if (pv != i) { const double t = sb[i]; sb[i] = sb[pv]; sb[pv] = t; }
        }
    __syncthreads();
    // forward: unit lower
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        if (tid < 32) {
            double x = tid < nb ? sb[k0 + tid] : 0.0;
            for (int j = 0; j < nb; ++j) {
                const double xj = __shfl_sync(0xffffffffu, x, j);
                if (tid > j && tid < nb) x -= LU[(size_t)(k0 + j) * n + k0 + tid] * xj;
            }
            if (tid < nb) sb[k0 + tid] = x;
        }
        __syncthreads();
        for (int r = k0 + nb + tid; r < n; r += 1024) {
            double acc = 0.0;
            for (int c = 0; c < nb; ++c) acc += LU[(size_t)(k0 + c) * n + r] * sb[k0 + c];
            sb[r] -= acc;
        }
        __syncthreads();
    }
    // backward: upper
    for (int k1 = n; k1 > 0; k1 -= NB) {
        const int k0 = max(0, k1 - NB), nb = k1 - k0;
        if (tid < 32) {
            double x = tid < nb ? sb[k0 + tid] : 0.0;
            for (int j = nb - 1; j >= 0; --j) {
                if (tid == j) x = x / LU[(size_t)(k0 + j) * n + k0 + j];
                const double xj = __shfl_sync(0xffffffffu, x, j);
                if (tid < j) x -= LU[(size_t)(k0 + j) * n + k0 + tid] * xj;
            }
            if (tid < nb) sb[k0 + tid] = x;
        }
        __syncthreads();
        for (int r = tid; r < k0; r += 1024) {
            double acc = 0.0;
            for (int c = 0; c < nb; ++c) acc += LU[(size_t)(k0 + c) * n + r] * sb[k0 + c];
            sb[r] -= acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += 1024) b[i] = sb[i];
}

// ---- getrs on a thread-block cluster: the single-CTA kernel above reads the whole factorisation through one SM
// and fetches every diagonal block with 32 dependent global loads.  Here 8 CTAs own the 32-row blocks of the
// right-hand side cyclically (block k belongs to CTA k % 8); per block step the owner's first warp solves the 32
// unknowns from registers (diagonal block loaded one step ahead, coalesced), pushes them into every peer's shared
// memory (DSMEM), and after ONE cluster barrier all CTAs update their own rows.  Arithmetic order is that of k_getrs.
constexpr int SOLVE_CTAS = 8, SOLVE_THREADS = 512, SOLVE_WARPS = SOLVE_THREADS / 32;
// acc = sum_c LU[(k0+c)*n + r] * xk[c] in ascending c, the loads issued 16 at a time ahead of the dependent sum
__device__ __forceinline__ double getrs_row_dot(const double* __restrict__ LU, int n, int k0, int nb, int r, const double* xk) {
    double acc = 0.0;
#pragma unroll
    for (int h = 0; h < NB; h += 16) {
        double a[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) a[c] = (h + c < nb) ? LU[(size_t)(k0 + h + c) * n + r] : 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (h + c < nb) acc += a[c] * xk[h + c];
    }
    return acc;
}
__global__ void __cluster_dims__(SOLVE_CTAS, 1, 1) __launch_bounds__(SOLVE_THREADS)
k_getrs_cluster(int n, const double* __restrict__ LU, const int* __restrict__ ipiv, double* __restrict__ b) {
    extern __shared__ double sb[];                   // n doubles; only this CTA's own rows are kept current
    __shared__ double xs[2][NB];                     // solved block of step s in xs[s & 1] (written by its owner)
    __shared__ double sd[NB][NB + 1];                // diagonal block of this CTA's next step: sd[j][i] = LU[k0+i, k0+j]
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int* sp = reinterpret_cast<int*>(sb + n);         // pivots staged in shared memory (the interchanges are serial)
    for (int i = tid; i < n; i += SOLVE_THREADS) { sb[i] = b[i]; sp[i] = ipiv[i] - 1; }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < n; ++i) {
            const int pv = sp[i];
            if (pv != i) { const double t = sb[i]; sb[i] = sb[pv]; sb[pv] = t; }
        }
    const int nblk = (n + NB - 1) / NB;
    // the diagonal block of step k is staged one step ahead by warp 1 of its owner (coalesced column loads)
    auto stage_diag = [&](int k) {
        const int k0 = k * NB;
#pragma unroll
        for (int h = 0; h < NB; h += 16) {
            double a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = (k0 + h + j < n && k0 + lane < n) ? LU[(size_t)(k0 + h + j) * n + k0 + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j) sd[h + j][lane] = a[j];
        }
    };
    auto push = [&](int s, int k0, int nb, double x) {
        if (lane < nb) {
            sb[k0 + lane] = x;
#pragma unroll
            for (int p = 0; p < SOLVE_CTAS; ++p) cl.map_shared_rank(&xs[s & 1][0], p)[lane] = x;
        }
    };
    int s = 0;
    // forward: unit lower
    if (rank == 0 && wid == 1) stage_diag(0);
    cl.sync();                                       // every CTA of the cluster is running before the first DSMEM push
    for (int k = 0; k < nblk; ++k, ++s) {
        const int k0 = k * NB, nb = min(NB, n - k0);
        const bool mine = rank == k % SOLVE_CTAS, next_mine = k + 1 < nblk && rank == (k + 1) % SOLVE_CTAS;
        if (mine && wid == 0) {
            double x = lane < nb ? sb[k0 + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const double xj = __shfl_sync(0xffffffffu, x, j);
                if (lane > j && lane < nb) x -= sd[j][lane] * xj;
            }
            push(s, k0, nb, x);
        }
        cl.sync();
        if (next_mine && wid == 1) stage_diag(k + 1);           // warp 0 of this CTA is past its use of sd (it is not the owner of k)
        const double* xk = xs[s & 1];
        int kb = k + 1 + ((rank - (k + 1)) % SOLVE_CTAS + SOLVE_CTAS) % SOLVE_CTAS;      // first own block after k
        // the owner of the next step updates that block first, with the warp that does not stage the diagonal
        for (kb += ((wid + SOLVE_WARPS - 2) % SOLVE_WARPS) * SOLVE_CTAS; kb < nblk; kb += SOLVE_WARPS * SOLVE_CTAS) {
            const int r = kb * NB + lane;
            if (r < n) sb[r] -= getrs_row_dot(LU, n, k0, nb, r, xk);
        }
        __syncthreads();
    }
    // backward: upper
    if (rank == (nblk - 1) % SOLVE_CTAS && wid == 1) stage_diag(nblk - 1);
    __syncthreads();
    for (int k = nblk - 1; k >= 0; --k, ++s) {
        const int k0 = k * NB, nb = min(NB, n - k0);
        const bool mine = rank == k % SOLVE_CTAS, next_mine = k >= 1 && rank == (k - 1) % SOLVE_CTAS;
        if (mine && wid == 0) {
            double x = lane < nb ? sb[k0 + lane] : 0.0;
#pragma unroll
            for (int j = NB - 1; j >= 0; --j) {
                if (j < nb) {
                    if (lane == j) x = x / sd[j][lane];
                    const double xj = __shfl_sync(0xffffffffu, x, j);
                    if (lane < j) x -= sd[j][lane] * xj;
                }
            }
            push(s, k0, nb, x);
        }
        cl.sync();
        if (next_mine && wid == 1) stage_diag(k - 1);
        const double* xk = xs[s & 1];
        // own blocks below k, the highest (the next steps' blocks) first
        const int nown = k > rank ? (k - rank + SOLVE_CTAS - 1) / SOLVE_CTAS : 0;        // own blocks kb = rank + q*CTAS < k
        for (int q = nown - 1 - (wid + SOLVE_WARPS - 2) % SOLVE_WARPS; q >= 0; q -= SOLVE_WARPS) {
            const int kb = rank + q * SOLVE_CTAS;
            const int r = kb * NB + lane;
            sb[r] -= getrs_row_dot(LU, n, k0, nb, r, xk);
        }
        __syncthreads();
    }
    for (int kb = rank; kb < nblk; kb += SOLVE_CTAS)
        for (int i = tid; i < NB; i += SOLVE_THREADS)
            if (kb * NB + i < n) b[kb * NB + i] = sb[kb * NB + i];
    cl.sync();                                       // no CTA leaves while a peer may still write its xs
}

// ---- y = C x in two deterministic stages
constexpr int MV_CH = 16;
__global__ void k_gemv_part(int n, const double* __restrict__ C, const double* __restrict__ x, double* __restrict__ part) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    const int per = (n + MV_CH - 1) / MV_CH;
    const int c0 = ch * per, c1 = min(n, c0 + per);
    if (r >= n) return;
    double acc = 0.0;
    for (int c = c0; c < c1; ++c) acc += C[(size_t)c * n + r] * x[c];
    part[(size_t)ch * n + r] = acc;
}
__global__ void k_gemv_sum(int n, const double* __restrict__ part, double* __restrict__ y) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double acc = 0.0;
    for (int ch = 0; ch < MV_CH; ++ch) acc += part[(size_t)ch * n + r];
    y[r] = acc;
}

}  // namespace ilm

using namespace ilm;

extern "C" int ilm_dense_factor(int n, double* A, int* ipiv, void* stream) {
    if (n < 0 || (n > 0 && (!A || !ipiv))) { set_error("ilm_dense_factor: bad arguments"); return ILM_EINVAL; }
    if (n == 0) return ILM_OK;
    DenseIo io(stream);
    double* dA = nullptr; int* dP = nullptr;
    ILM_TRY(io.map(A, (size_t)n * n, true, true, &dA));
    ILM_TRY(io.map(ipiv, (size_t)n, false, true, &dP));
    cudaStream_t st = io.st;
    ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    static const bool panel_smem_only = getenv("ILM_LU_PANEL_SMEM") != nullptr;     // the shared-memory panel, kept for comparison
    // outer block (multiple of NB; ILM_LU_OUTER).  KB = NB is the plain right-looking algorithm and the default:
    // measured on B200 after the trailing update stopped serialising its read-modify-write of C, rank-32 updates
    // beat the two-level form at every size tried (N = 4593: 23.9 vs 25.2 ms with KB = 128; N = 2295: 8.6 vs 10.6 ms)
    // -- the inner trsm / gemm / swap launches on a <= 128-column block cost more than the saved sweeps.
    static const int kb_env = getenv("ILM_LU_OUTER") ? atoi(getenv("ILM_LU_OUTER")) : 0;
    const int KB = kb_env >= NB ? (kb_env / NB) * NB : NB;
    ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_regs, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_rot, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    auto panel = [&](int k0, int nb) -> int {
        if (!panel_smem_only && n - k0 <= PR_MAXC * PR_THREADS) {
            ILM_TRY(launch_panel_regs(n, k0, nb, dA, dP, st));
        } else {
            const int chunk = (n - k0 + PANEL_CTAS - 1) / PANEL_CTAS;
            const int ldp = chunk | 1;
            const size_t smem = (size_t)ldp * NB * sizeof(double);
            if (smem <= 200 * 1024) {
                k_lu_panel_smem<<<PANEL_CTAS, 1024, smem, st>>>(n, k0, nb, ldp, dA, dP);
            } else if (n - k0 >= 2048) {
                k_lu_panel_cluster<<<PANEL_CTAS, 1024, 0, st>>>(n, k0, nb, dA, dP);
            } else {
                k_lu_panel<<<1, 1024, 0, st>>>(n, k0, nb, dA, dP);
            }
        }
        g_dense_launches++;
        return ILM_OK;
    };
    // C[M x Nc] -= A[M x K] * B[K x Nc], all inside dA (leading dimension n)
    auto gemm = [&](int M, int Nc, int K, int arow, int acol, int brow, int bcol) {
        if (M <= 0 || Nc <= 0 || K <= 0) return;
        dim3 grid((M + 63) / 64, (Nc + 63) / 64);
        k_lu_gemm<<<grid, 128, 0, st>>>(M, Nc, K, dA + (size_t)acol * n + arow, dA + (size_t)bcol * n + brow,
                                        dA + (size_t)bcol * n + arow, n);
        g_dense_launches++;
    };
    auto trsm = [&](int k0, int nb, int cbeg, int cend) {
        if (cend <= cbeg) return;
        k_lu_trsm<<<(cend - cbeg + 63) / 64, 64, 0, st>>>(n, k0, nb, dA, cbeg, cend);
        g_dense_launches++;
    };
    // Right-looking LU with LOOK-AHEAD (the default, KB == NB): the panel factorisation is latency bound (a cluster of
    // 8-16 SMs, one cluster barrier per column) and used to leave the other ~130 SMs idle for more than half of the
    // factorisation.  Two internal streams: `hi` (high priority) carries panel k+1 and the narrow update that feeds it
    // (interchanges, triangular solve and rank-32 update of the next 32 columns only), `lo` carries the wide update of
    // step k (all other columns, left and right interchanges) and runs underneath the next panel.
    static const bool no_lookahead = getenv("ILM_LU_NO_LOOKAHEAD") != nullptr;
    if (KB == NB && !no_lookahead && n > 2 * NB) {
        struct Aux {
            cudaStream_t hi = nullptr, lo = nullptr;
            cudaEvent_t begin = nullptr, evP = nullptr, evW = nullptr, end = nullptr, evW2[2] = {nullptr, nullptr};
            int* perm = nullptr; size_t perm_cap = 0;               // net permutations of the panels (k_lu_panel_v2 -> k_lu_swap_trsm)
            double* uscr = nullptr;                                 // 2 x (32 x 32) block rows in flight inside the panel kernels
        };
        static Aux aux_dev[64];
        int dev = 0;
        ILM_CUDA(cudaGetDevice(&dev));
        Aux& ax = aux_dev[dev & 63];
        if (!ax.hi) {
            int lo_p = 0, hi_p = 0;
            ILM_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            ILM_CUDA(cudaStreamCreateWithPriority(&ax.hi, cudaStreamNonBlocking, hi_p));
            ILM_CUDA(cudaStreamCreateWithPriority(&ax.lo, cudaStreamNonBlocking, lo_p));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.begin, cudaEventDisableTiming));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.evP, cudaEventDisableTiming));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.evW, cudaEventDisableTiming));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.end, cudaEventDisableTiming));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.evW2[0], cudaEventDisableTiming));
            ILM_CUDA(cudaEventCreateWithFlags(&ax.evW2[1], cudaEventDisableTiming));
            ILM_CUDA(cudaMalloc(&ax.uscr, 2 * NB * NB * sizeof(double)));
        }
        const cudaStream_t user = st;
        ILM_CUDA(cudaEventRecord(ax.begin, user));
        ILM_CUDA(cudaStreamWaitEvent(ax.hi, ax.begin, 0));
        ILM_CUDA(cudaStreamWaitEvent(ax.lo, ax.begin, 0));
        bool wide_pending = false;
        static const bool v1 = getenv("ILM_LU_V1") != nullptr;        // the separate narrow-update launches (first round-2 form)
        if (!v1 && n <= P2_MAXC * 512) {
            // Second form: the panel kernel applies the previous panel's update to its own 32 columns (k_lu_panel_v2);
            // the wide update is two launches (net permutation + block row, rank-32 DMMA update) on the low-priority stream.
            ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_v2<256, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_v2<512, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_v2<256, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            ILM_CUDA(cudaFuncSetAttribute(k_lu_panel_v2<512, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            const int npan = (n + NB - 1) / NB;
            if (ax.perm_cap < (size_t)npan * PERM_STRIDE) {
                ILM_CUDA(cudaStreamSynchronize(ax.hi));
                ILM_CUDA(cudaStreamSynchronize(ax.lo));
                cudaFree(ax.perm);
                ax.perm = nullptr; ax.perm_cap = 0;
                ILM_CUDA(cudaMalloc(&ax.perm, (size_t)npan * PERM_STRIDE * sizeof(int)));
                ax.perm_cap = (size_t)npan * PERM_STRIDE;
            }
            static const bool want_trace = getenv("ILM_LU_TRACE") != nullptr;
            long long* trace = nullptr;
            const size_t ntrace = 16 + 2 * (size_t)npan;
            if (want_trace) { ILM_CUDA(cudaMalloc(&trace, ntrace * sizeof(long long))); ILM_CUDA(cudaMemsetAsync(trace, 0, ntrace * sizeof(long long), ax.hi)); }
            int pk0 = 0, pnb = 0;
            for (int k0 = 0, ip = 0; k0 < n; k0 += NB, ++ip) {
                const int nb = n - k0 < NB ? n - k0 : NB, k1 = k0 + nb;
                const int nb1 = n - k1 < NB ? n - k1 : NB, k1e = k1 + nb1;      // the next panel's columns [k1, k1e)
                int* perm = ax.perm + (size_t)ip * PERM_STRIDE;
                // this panel's columns were last touched by the wide update of step k-2 (step k-1 excluded them: they get that
                // update inside the panel kernel), so the panel runs beside the wide update of step k-1
                if (ip >= 2) ILM_CUDA(cudaStreamWaitEvent(ax.hi, ax.evW2[ip & 1], 0));
                ILM_TRY(launch_panel_v2(n, k0, nb, dA, dP, pk0, pnb, perm, ax.uscr + (size_t)(ip & 1) * NB * NB, trace, ax.hi));
                g_dense_launches++;
                ILM_CUDA(cudaEventRecord(ax.evP, ax.hi));
                ILM_CUDA(cudaStreamWaitEvent(ax.lo, ax.evP, 0));
                const int ncols = k0 + (n - k1e);
                if (ncols > 0) {
                    k_lu_swap_trsm<<<(ncols + 63) / 64, 64, 0, ax.lo>>>(n, k0, nb, dA, perm, k1e);
                    g_dense_launches++;
                }
                st = ax.lo;
                if (n - k1e > 0) gemm(n - k1, n - k1e, nb, k1, k0, k0, k1e);
                ILM_CUDA(cudaEventRecord(ax.evW2[ip & 1], ax.lo));
                ILM_CUDA(cudaEventRecord(ax.evW, ax.lo));
                wide_pending = true;
                pk0 = k0; pnb = nb;
            }
            ILM_CUDA(cudaStreamWaitEvent(ax.hi, ax.evW, 0));
            ILM_CUDA(cudaEventRecord(ax.end, ax.hi));
            ILM_CUDA(cudaStreamWaitEvent(user, ax.end, 0));
            st = user;
            ILM_CUDA(cudaGetLastError());
            if (trace) {
                std::vector<long long> hv(ntrace);
                long long* h = hv.data();
                ILM_CUDA(cudaStreamSynchronize(user));
                ILM_CUDA(cudaMemcpy(h, trace, ntrace * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(trace);
                {
                    double dur = 0, gap = 0;
                    for (int i = 0; i < npan; ++i) dur += (double)(h[17 + 2 * i] - h[16 + 2 * i]);
                    for (int i = 1; i < npan; ++i) gap += (double)(h[16 + 2 * i] - h[15 + 2 * i]);
                    fprintf(stderr, "ilm LU timeline: panel kernels %.1f us each, gap between panel kernels %.1f us, first entry to last exit %.3f ms\n",
                            dur / npan * 1e-3, npan > 1 ? gap / (npan - 1) * 1e-3 : 0.0, (double)(h[17 + 2 * (npan - 1)] - h[16]) * 1e-6);
                    for (int i = 0; i < npan; i += npan / 8 > 0 ? npan / 8 : 1)
                        fprintf(stderr, "   panel %3d: %.1f us, gap before %.1f us\n", i, (double)(h[17 + 2 * i] - h[16 + 2 * i]) * 1e-3,
                                i ? (double)(h[16 + 2 * i] - h[15 + 2 * i]) * 1e-3 : 0.0);
                }
                static const char* nm[11] = {"(stamp slot 0)", "argmax warp", "keys -> smem", "block barrier", "argmax CTA + row -> smem", "barrier + push",
                                             "cluster barrier", "argmax cluster", "update", "rotations (per panel)", "write-back + perm (per panel)"};
                fprintf(stderr, "ilm LU trace, n = %d, %d panels (clock cycles seen by thread 0 of CTA 0):\n", n, npan);
                fprintf(stderr, "  narrow update: U block row (warp 0) + row loads %.0f, wait for U %.0f, push + cluster barrier %.0f, rank-32 update %.0f cycles per panel\n",
                        (double)h[11] / npan, (double)h[12] / npan, (double)h[13] / npan, (double)h[14] / npan);
                for (int i = 0; i < 11; ++i) {
                    const bool per_panel = i == 0 || i >= 9;
                    fprintf(stderr, "  %-32s %10.1f cycles %s\n", nm[i], (double)h[i] / (per_panel ? npan : n), per_panel ? "per panel" : "per column");
                }
            }
            return io.finish();
        }
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int nb = n - k0 < NB ? n - k0 : NB, k1 = k0 + nb;
            const int nb1 = n - k1 < NB ? n - k1 : NB, k1e = k1 + nb1;          // the next panel's columns [k1, k1e)
            st = ax.hi;
            ILM_TRY(panel(k0, nb));                                             // runs beside the wide update of step k-1
            ILM_CUDA(cudaEventRecord(ax.evP, ax.hi));
            ILM_CUDA(cudaStreamWaitEvent(ax.lo, ax.evP, 0));
            if (wide_pending) ILM_CUDA(cudaStreamWaitEvent(ax.hi, ax.evW, 0)); // the next columns carry every earlier update
            if (nb1 > 0) {                                                      // narrow update: only what panel k+1 needs
                ILM_TRY(launch_swap(n, k0, nb, dA, dP, k1, k1e, k1e, k1e, st));
                trsm(k0, nb, k1, k1e);
                gemm(n - k1, nb1, nb, k1, k0, k0, k1);
            }
            st = ax.lo;                                                         // wide update: everything else
            ILM_TRY(launch_swap(n, k0, nb, dA, dP, 0, n, k0, k1e, st));
            if (n - k1e > 0) {
                trsm(k0, nb, k1e, n);
                gemm(n - k1, n - k1e, nb, k1, k0, k0, k1e);
            }
            ILM_CUDA(cudaEventRecord(ax.evW, ax.lo));
            wide_pending = true;
        }
        ILM_CUDA(cudaStreamWaitEvent(ax.hi, ax.evW, 0));
        ILM_CUDA(cudaEventRecord(ax.end, ax.hi));
        ILM_CUDA(cudaStreamWaitEvent(user, ax.end, 0));
        st = user;
        ILM_CUDA(cudaGetLastError());
        return io.finish();
    }
    // Two-level right-looking LU: 32-wide register panels inside an outer block of KB columns; the inner trailing
    // updates touch only the outer block, the rest of the matrix gets ONE rank-KB update per outer block (the matrix
    // is read and written n/KB times instead of n/32 times).
    for (int K0 = 0; K0 < n; K0 += KB) {
        const int kb = n - K0 < KB ? n - K0 : KB, Kend = K0 + kb;
        for (int k0 = K0; k0 < Kend; k0 += NB) {
            const int nb = Kend - k0 < NB ? Kend - k0 : NB;
            ILM_TRY(panel(k0, nb));
            ILM_TRY(launch_swap(n, k0, nb, dA, dP, K0, Kend, k0, k0 + nb, st));
            trsm(k0, nb, k0 + nb, Kend);
            gemm(n - (k0 + nb), Kend - (k0 + nb), nb, k0 + nb, k0, k0, k0 + nb);
        }
        ILM_TRY(launch_swap(n, K0, kb, dA, dP, 0, n, K0, Kend, st));
        if (n - Kend > 0) {
            for (int k0 = K0; k0 < Kend; k0 += NB) {       // block row: A12 <- L11^-1 A12
                const int nb = Kend - k0 < NB ? Kend - k0 : NB;
                trsm(k0, nb, Kend, n);
                gemm(Kend - (k0 + nb), n - Kend, nb, k0 + nb, k0, k0, Kend);
            }
            gemm(n - Kend, n - Kend, kb, Kend, K0, K0, Kend);
        }
    }
    ILM_CUDA(cudaGetLastError());
    return io.finish();
}

extern "C" int ilm_dense_solve(int n, const double* LU, const int* ipiv, int nrhs, double* B, void* stream) {
    if (n < 0 || nrhs < 0 || (n > 0 && nrhs > 0 && (!LU || !ipiv || !B))) { set_error("ilm_dense_solve: bad arguments"); return ILM_EINVAL; }
    if (n == 0 || nrhs == 0) return ILM_OK;
    if ((size_t)n * (sizeof(double) + sizeof(int)) > 200 * 1024) { set_error("ilm_dense_solve: n too large for the shared-memory solve"); return ILM_ESIZE; }
    DenseIo io(stream);
    double* dLU = nullptr; int* dP = nullptr; double* dB = nullptr;
    ILM_TRY(io.map(const_cast<double*>(LU), (size_t)n * n, true, false, &dLU));
    ILM_TRY(io.map(const_cast<int*>(ipiv), (size_t)n, true, false, &dP));
    ILM_TRY(io.map(B, (size_t)n * nrhs, true, true, &dB));
    const size_t smem = (size_t)n * sizeof(double);
    static const bool single_cta = getenv("ILM_GETRS_SINGLE_CTA") != nullptr;      // the first version, kept for comparison
    if (n < 4 * NB || single_cta) {
        ILM_CUDA(cudaFuncSetAttribute(k_getrs, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int r = 0; r < nrhs; ++r) k_getrs<<<1, 1024, smem, io.st>>>(n, dLU, dP, dB + (size_t)r * n);
    } else {
        ILM_CUDA(cudaFuncSetAttribute(k_getrs_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int r = 0; r < nrhs; ++r)
            k_getrs_cluster<<<SOLVE_CTAS, SOLVE_THREADS, smem + (size_t)n * sizeof(int), io.st>>>(n, dLU, dP, dB + (size_t)r * n);
    }
    g_dense_launches += nrhs;
    ILM_CUDA(cudaGetLastError());
    return io.finish();
}

extern "C" int ilm_dense_matvec_pow(int n, const double* C, int k, double* s, void* stream) {
    if (n < 0 || k < 0 || (n > 0 && (!C || !s))) { set_error("ilm_dense_matvec_pow: bad arguments"); return ILM_EINVAL; }
    if (n == 0 || k == 0) return ILM_OK;
    DenseIo io(stream);
    double* dC = nullptr; double* ds = nullptr; double* part = nullptr;
    ILM_TRY(io.map(const_cast<double*>(C), (size_t)n * n, true, false, &dC));
    ILM_TRY(io.map(s, (size_t)n, true, true, &ds));
    ILM_CUDA(cudaMalloc(&part, (size_t)MV_CH * n * sizeof(double)));
    io.owned.push_back(part);
    for (int it = 0; it < k; ++it) {
        k_gemv_part<<<dim3((n + 127) / 128, MV_CH), 128, 0, io.st>>>(n, dC, ds, part);
        k_gemv_sum<<<(n + 127) / 128, 128, 0, io.st>>>(n, part, ds);
        g_dense_launches += 2;
    }
    ILM_CUDA(cudaGetLastError());
    int s_ = io.finish();
    ILM_CUDA(cudaStreamSynchronize(io.st));     // `part` is freed by ~DenseIo
    return s_;
}

extern "C" int64_t ilm_dense_launch_count(void) { return (int64_t)ilm::g_dense_launches; }

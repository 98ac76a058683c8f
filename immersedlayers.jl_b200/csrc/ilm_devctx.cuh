// ilm_devctx.cuh -- device execution context of the convolution kernel bodies (named barriers,
// bulk-tensor copies, L2 prefetch); the CPU emulation harness supplies its own (test_fft_host.cu).
#pragma once
#include <cuda.h>

#include "ilm_internal.h"

namespace ilm {

struct DevCtx {
    int tid, grp;
    const void* tmap;        // CUtensorMap of the S2 spectrum (pass C), else null
    ILM_HD void sync() {
#ifdef __CUDA_ARCH__
        asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
#endif
    }
    // true if the predicate holds for any thread of the calling warp
    ILM_HD bool any(bool pred) {
#ifdef __CUDA_ARCH__
        return __any_sync(0xffffffffu, pred) != 0;
#else
        return pred;
#endif
    }
    ILM_HD void sync_cta() {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    // producer/consumer hand-off between the two groups (512 = both groups)
    ILM_HD void arrive(int id) {
#ifdef __CUDA_ARCH__
        __threadfence_block();
        asm volatile("bar.arrive %0, 512;" ::"r"(id) : "memory");
#endif
    }
    ILM_HD void wait(int id) {
#ifdef __CUDA_ARCH__
        asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory");
#endif
    }
    // thread-block cluster (big column pass): rank of this CTA and a barrier over the whole cluster that
    // also orders the global-memory hand-off line (release / acquire at cluster scope)
    ILM_HD int cluster_rank() {
#ifdef __CUDA_ARCH__
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        return (int)r;
#else
        return 0;
#endif
    }
    ILM_HD void cluster_sync() {
#ifdef __CUDA_ARCH__
        __threadfence();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
    }
    ILM_HD void delay(int ns) {
#ifdef __CUDA_ARCH__
        if (ns > 0) __nanosleep((unsigned)ns);
#endif
    }
    // ---- bulk-tensor (TMA) row loads of pass C -------------------------------------------
    ILM_HD void tma_init(double2* mbar) {
#ifdef __CUDA_ARCH__
        if (tid == 0) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        sync();
#endif
    }
    // rows row0 .. row0+nrows-1 of parity px (L complex each, m order) into dst + r*slot_stride
    ILM_HD void tma_load_rows(const ConvArgs& a, int px, int row0, int nrows, int L, int slot_stride, double2* dst,
                              double2* mbar) {
#ifdef __CUDA_ARCH__
        if (tid == 0) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
            const unsigned long long tm = reinterpret_cast<unsigned long long>(tmap);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(nrows * L * 16) : "memory");
            if (a.s2_rowmajor) {                                    // one contiguous bulk copy per row (probe path)
                for (int r = 0; r < nrows; ++r) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * slot_stride);
                    const double2* src = a.S2 + s2rm_index(a, px, 0, row0 + r);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(d), "l"(src), "r"(L * 16), "r"(mb) : "memory");
                }
                return;
            }
            const int na = L >> 1;                                 // tile columns per row
            const int box = na < 256 ? na : 256;
            for (int r = 0; r < nrows; ++r) {
                const int row = s_row(a.g, row0 + r);
                for (int a0 = 0; a0 < na; a0 += box) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * slot_stride + 2 * a0);
                    const int c0 = (row & 1) * 4, c1 = row >> 1, c2 = px * na + a0;
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(d), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(mb) : "memory");
                }
            }
        }
#endif
    }
    ILM_HD void tma_wait(double2* mbar, int phase) {
#ifdef __CUDA_ARCH__
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(mb), "r"(phase) : "memory");
#endif
    }
    ILM_HD void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
    }
};

}  // namespace ilm

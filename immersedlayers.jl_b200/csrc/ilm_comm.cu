// ilm_comm.cu -- NCCL behind the C ABI: one process per GPU, the communicator is bound to
// the plan and every collective of the sharded entry points runs on the plan's stream.
//
// The reference is a single Julia process (no distributed path); what shards naturally is the column loop of the
// Schur builders (src/matrix_operators.jl:16-26: N independent probes) -- each rank probes a contiguous block of
// columns into its slice of the device-resident matrix and one grouped in-place broadcast (an all-gather with
// uneven counts) completes S on every rank.
//
// NCCL is resolved with dlopen/dlsym at ilm_comm_init time, not linked: the library keeps loading (and every
// single-GPU entry point keeps working) on a box without NCCL, and inside a torch process the already-loaded
// libnccl.so.2 of torch is the one that is used.  Failures map to ILM_ENCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "ilm_internal.h"

namespace ilm {

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
};

NcclApi* nccl() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a.ok ? &a : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy the host process already uses (torch)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error(std::string("NCCL not found: ") + dlerror()); return nullptr; }
#define ILM_SYM(name)                                                                   \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name));                \
    if (!a.name) { set_error("NCCL symbol nccl" #name " not found"); return nullptr; }
    ILM_SYM(GetUniqueId) ILM_SYM(CommInitRank) ILM_SYM(CommDestroy) ILM_SYM(GroupStart) ILM_SYM(GroupEnd)
    ILM_SYM(Broadcast) ILM_SYM(Send) ILM_SYM(Recv) ILM_SYM(AllReduce) ILM_SYM(GetErrorString)
#undef ILM_SYM
    a.ok = true;
    return &a;
}

int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
    set_error(std::string(what) + ": " + (a && a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
    return ILM_ENCCL;
}

}  // namespace

#define ILM_NCCL(call)                                                        \
    do {                                                                      \
        ncclResult_t r__ = (call);                                            \
        if (r__ != ncclSuccess) return nccl_fail(api, r__, #call);            \
    } while (0)

// contiguous split of n columns into even-sized blocks, so that the column PAIRS that share one complex transform
// stay on one rank (the same rule as shard.column_ranges)
void comm_column_range(int n, int nranks, int rank, int* lo, int* hi) {
    const int pairs = (n + 1) / 2, base = pairs / nranks, rem = pairs % nranks;
    const int start = rank * base + (rank < rem ? rank : rem), cnt = base + (rank < rem ? 1 : 0);
    *lo = 2 * start < n ? 2 * start : n;
    *hi = 2 * (start + cnt) < n ? 2 * (start + cnt) : n;
}

// in-place all-gather of the column blocks of a device-resident column-major matrix with `ld` rows
int comm_allgather_columns(ilm_plan* p, double* dA, int ld, int ncols) {
    NcclApi* api = nccl();
    if (!api || !p->comm) { set_error("the plan has no communicator (ilm_comm_init)"); return ILM_ENCCL; }
    ncclComm_t comm = static_cast<ncclComm_t>(p->comm);
    ILM_NCCL(api->GroupStart());
    for (int r = 0; r < p->comm_size; ++r) {
        int lo, hi;
        comm_column_range(ncols, p->comm_size, r, &lo, &hi);
        if (hi > lo) {
            double* blk = dA + (size_t)lo * ld;
            ILM_NCCL(api->Broadcast(blk, blk, (size_t)(hi - lo) * ld, ncclDouble, r, comm, p->stream));
        }
    }
    ILM_NCCL(api->GroupEnd());
    return ILM_OK;
}

// the same with explicit column boundaries: rank r owns the columns [bounds[r], bounds[r + 1])
int comm_allgather_column_blocks(ilm_plan* p, double* dA, int ld, const int* bounds) {
    NcclApi* api = nccl();
    if (!api || !p->comm) { set_error("the plan has no communicator (ilm_comm_init)"); return ILM_ENCCL; }
    ncclComm_t comm = static_cast<ncclComm_t>(p->comm);
    ILM_NCCL(api->GroupStart());
    for (int r = 0; r < p->comm_size; ++r) {
        if (bounds[r + 1] > bounds[r]) {
            double* blk = dA + (size_t)bounds[r] * ld;
            ILM_NCCL(api->Broadcast(blk, blk, (size_t)(bounds[r + 1] - bounds[r]) * ld, ncclDouble, r, comm, p->stream));
        }
    }
    ILM_NCCL(api->GroupEnd());
    return ILM_OK;
}

// variable-count all-to-all on device buffers (counts in doubles, peer-major packing): the exchange of the slab solve
int comm_alltoallv(ilm_plan* p, const double* send, const int64_t* scount, double* recv, const int64_t* rcount) {
    NcclApi* api = nccl();
    if (!api || !p->comm) { set_error("the plan has no communicator (ilm_comm_init)"); return ILM_ENCCL; }
    ncclComm_t comm = static_cast<ncclComm_t>(p->comm);
    ILM_NCCL(api->GroupStart());
    size_t so = 0, ro = 0;
    for (int r = 0; r < p->comm_size; ++r) {
        if (scount[r] > 0) ILM_NCCL(api->Send(send + so, (size_t)scount[r], ncclDouble, r, comm, p->stream));
        if (rcount[r] > 0) ILM_NCCL(api->Recv(recv + ro, (size_t)rcount[r], ncclDouble, r, comm, p->stream));
        so += (size_t)scount[r];
        ro += (size_t)rcount[r];
    }
    ILM_NCCL(api->GroupEnd());
    return ILM_OK;
}

void comm_release(ilm_plan* p) {
    if (p->comm) {
        NcclApi* api = nccl();
        if (api) api->CommDestroy(static_cast<ncclComm_t>(p->comm));
        p->comm = nullptr;
    }
    p->comm_rank = 0;
    p->comm_size = 1;
}

}  // namespace ilm

using namespace ilm;

extern "C" int ilm_comm_unique_id(void* id, int nbytes) {
    if (!id || nbytes < (int)sizeof(ncclUniqueId)) { set_error("ilm_comm_unique_id: need a 128-byte buffer"); return ILM_EINVAL; }
    NcclApi* api = nccl();
    if (!api) return ILM_ENCCL;
    ncclUniqueId u;
    ILM_NCCL(api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return ILM_OK;
}

extern "C" int ilm_comm_init(ilm_plan* p, const void* id, int nbytes, int rank, int nranks) {
    if (!p || !id || nbytes < (int)sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("ilm_comm_init: bad arguments");
        return ILM_EINVAL;
    }
    if (p->shared) { set_error("ilm_comm_init: bind the communicator to the parent plan"); return ILM_EINVAL; }
    cudaSetDevice(p->device);
    NcclApi* api = nccl();
    if (!api) return ILM_ENCCL;
    comm_release(p);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    ILM_NCCL(api->CommInitRank(&comm, nranks, u, rank));
    p->comm = comm;
    p->comm_rank = rank;
    p->comm_size = nranks;
    return ILM_OK;
}

extern "C" int ilm_comm_destroy(ilm_plan* p) {
    if (!p) { set_error("null plan"); return ILM_EINVAL; }
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    comm_release(p);
    return ILM_OK;
}

extern "C" int ilm_comm_info(const ilm_plan* p, int* rank, int* nranks) {
    if (!p) { set_error("null plan"); return ILM_EINVAL; }
    if (rank) *rank = p->comm_rank;
    if (nranks) *nranks = p->comm_size;
    return ILM_OK;
}

extern "C" int ilm_column_range(int n, int nranks, int rank, int* lo, int* hi) {
    if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) { set_error("ilm_column_range: bad arguments"); return ILM_EINVAL; }
    comm_column_range(n, nranks, rank, lo, hi);
    return ILM_OK;
}

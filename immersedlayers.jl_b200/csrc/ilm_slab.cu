// ilm_slab.cu -- slab decomposition of ONE lattice-Green's-function convolution over several
// GPUs (BASELINE config C5b: the largest single grid).
//
// The reference has no distributed path (one Julia process, FFTW threads); this is the multi-GPU
// form of inverse_laplacian! (src/grid_operators.jl:153-179) for a field whose rows are spread
// over `nranks` GPUs, one process per GPU:
//
//   rows [row0,row1) of the field --pass A (FFT_x)--> spectrum rows of this rank
//        == all-to-all #1 ==>  all rows of the x-frequency columns [tc0,tc1) this rank owns
//        --pass B (FFT_y * Ghat * IFFT_y)-->
//        == all-to-all #2 ==>  this rank's rows of every column
//        --pass C (IFFT_x)--> rows [row0,row1) of the result.
//
// The library does the compute and the pack/unpack of the exchange buffers; the collective itself
// is issued by the host language (torch.distributed all_to_all_single over NCCL in shard.py,
// NCCL.jl / MPI in a Julia shim) between ilm_slab_forward / _columns / _inverse.  The partition is a
// pure function of (grid, nranks, rank), so every rank computes every peer's counts locally.
//
// Spectrum layout (ilm_conv.cuh): S[tile column][row pair][2x2 tile]; a tile column (two
// x-frequency columns) is a contiguous run of 2*MYp complex numbers, and a rank's row range is a
// contiguous piece of it.  Two consequences used here:
//   * with the COMPACT geometry "MYp = this rank's row count" the spectrum of this rank's rows is
//     [tile column][my row pairs], and the tile columns of peer q are a contiguous block of it: pass A
//     writes the send buffer of exchange #1 directly and pass C reads the receive buffer of exchange #2
//     directly (no pack, no unpack, no full-size buffer on either side);
//   * the column pass needs [my tile columns][all rows]: the only copies left are the unpack of
//     exchange #1 and the pack of exchange #2 (one strided-2D copy per peer), into / out of buffers that
//     hold this rank's tile columns only (1/nranks of the full spectrum).
#include <algorithm>

#include "ilm_internal.h"

using namespace ilm;

namespace {

struct Range { int lo, hi; };

// contiguous split of n units over nranks
Range split(int n, int nranks, int r) {
    const int base = n / nranks, rem = n % nranks;
    const int lo = r * base + std::min(r, rem);
    return {lo, lo + base + (r < rem ? 1 : 0)};
}

int fill_info(int NX, int NY, int rows, int nranks, int rank, ilm_slab_info* o) {
    if (!o || nranks < 1 || rank < 0 || rank >= nranks || NX < 2 || NY < 2 || rows < 1 || rows > NY + 1) {
        set_error("ilm_slab_partition: bad arguments");
        return ILM_EINVAL;
    }
    o->nranks = nranks; o->rank = rank;
    o->Lx = conv_half_len(NX); o->Ly = conv_half_len(NY);
    o->MYp = (rows + 1) & ~1;
    const Range rp = split(o->MYp / 2, nranks, rank);          // row pairs: a 2x2 tile is never split
    o->row0 = 2 * rp.lo; o->row1 = 2 * rp.hi;
    const int F = 4096 / o->Ly > 0 ? 4096 / o->Ly : 1;          // FFTs per group of the column pass
    o->cpw = (F <= 1) ? 2 : F;                                  // x-frequency columns per pass-B work item
    const int nwork = (2 * o->Lx + o->cpw - 1) / o->cpw;
    const Range wp = split(nwork, nranks, rank);
    o->wlo = wp.lo; o->whi = wp.hi;
    o->ntc = o->Lx;                                             // tile columns in total (both parities)
    o->tc0 = std::min(wp.lo * o->cpw / 2, o->ntc);
    o->tc1 = std::min(wp.hi * o->cpw / 2, o->ntc);
    return ILM_OK;
}

struct Peer { int nrows, row0, ntc, tc0; };
Peer peer_of(const ilm_slab_info* me, int NX, int NY, int r) {
    ilm_slab_info t{};
    fill_info(NX, NY, me->MYp, me->nranks, r, &t);             // rows = MYp gives the same MYp
    return {t.row1 - t.row0, t.row0, t.tc1 - t.tc0, t.tc0};
}

// blocks are (tile columns) x (2 * rows) complex; pitch of the spectrum = 2*MYp complex
int copy_block(ilm_plan* p, double2* spec, int MYp, int tc0, int ntc, int row0, int nrows, double2* buf, bool pack) {
    if (ntc == 0 || nrows == 0) return ILM_OK;
    const size_t pitch = (size_t)2 * MYp * sizeof(double2), width = (size_t)2 * nrows * sizeof(double2);
    double2* s = spec + (size_t)tc0 * 2 * MYp + (size_t)2 * row0;
    if (pack) ILM_CUDA(cudaMemcpy2DAsync(buf, width, s, pitch, width, ntc, cudaMemcpyDeviceToDevice, p->stream));
    else ILM_CUDA(cudaMemcpy2DAsync(s, pitch, buf, width, width, ntc, cudaMemcpyDeviceToDevice, p->stream));
    return ILM_OK;
}

int check_info(const ilm_plan* p, const ilm_slab_info* s, int my1, int my2) {
    ilm_slab_info t{};
    const int rows = std::max(my1, my2);
    if (!s || fill_info(p->g.NX, p->g.NY, rows, s->nranks, s->rank, &t) != ILM_OK || t.MYp != s->MYp || t.row0 != s->row0 ||
        t.row1 != s->row1 || t.tc0 != s->tc0 || t.tc1 != s->tc1 || t.Lx != p->Lx || t.Ly != p->Ly) {
        set_error("ilm_slab: the slab info does not belong to this plan / these layouts");
        return ILM_EINVAL;
    }
    return ILM_OK;
}

ConvArgs base_args(ilm_plan* p, const ilm_slab_info* s, int kernel_id) {
    ConvArgs a = conv_base_args(p);
    a.g = ConvGeom{p->Lx, p->Ly, s->MYp, s->MYp};
    a.Ghat = p->kernels[kernel_id].ghat;
    a.rlo = 0; a.rhi = s->MYp; a.olo = 0; a.ohi = s->MYp;
    return a;
}

// FieldRef in LOCAL rows: row 0 is the rank's first row; my = how many of the rank's rows exist in this layout
FieldRef local_ref(const ilm_plan* p, int layout, double* rows_ptr, int row0, int nrows) {
    if (!rows_ptr || layout < 0) return FieldRef{nullptr, 0, 0};
    const LayoutInfo li = layout_info(layout, p->g.NX, p->g.NY);
    int my = li.my - row0;
    my = my < 0 ? 0 : (my > nrows ? nrows : my);
    return FieldRef{rows_ptr, li.mx, my};
}

int layout_rows(const ilm_plan* p, int layout) { return layout < 0 ? 0 : layout_info(layout, p->g.NX, p->g.NY).my; }
bool scalar_layout(int l) { return l >= ILM_NODES_PRIMAL && l <= ILM_YEDGES; }

}  // namespace

extern "C" int ilm_slab_partition(int NX, int NY, int rows, int nranks, int rank, ilm_slab_info* out) {
    return fill_info(NX, NY, rows, nranks, rank, out);
}

extern "C" int ilm_slab_counts(int NX, int NY, const ilm_slab_info* me, int phase, int64_t* send, int64_t* recv) {
    if (!me || !send || !recv || (phase != 0 && phase != 1)) { set_error("ilm_slab_counts: bad arguments"); return ILM_EINVAL; }
    const Peer self = peer_of(me, NX, NY, me->rank);
    for (int r = 0; r < me->nranks; ++r) {
        const Peer q = peer_of(me, NX, NY, r);
        const int64_t mine_rows_their_cols = (int64_t)4 * q.ntc * self.nrows;      // doubles
        const int64_t their_rows_mine_cols = (int64_t)4 * self.ntc * q.nrows;
        send[r] = phase == 0 ? mine_rows_their_cols : their_rows_mine_cols;
        recv[r] = phase == 0 ? their_rows_mine_cols : mine_rows_their_cols;
    }
    return ILM_OK;
}

extern "C" int64_t ilm_slab_buffer_doubles(const ilm_slab_info* me) {
    if (!me) return 0;
    const int64_t a = (int64_t)4 * me->ntc * (me->row1 - me->row0), b = (int64_t)4 * (me->tc1 - me->tc0) * me->MYp;
    return std::max<int64_t>(std::max(a, b), 1);
}

#define ILM_SLAB_PLAN(p)                                        \
    do {                                                        \
        if (!(p)) { set_error("null plan"); return ILM_EINVAL; } \
        cudaSetDevice((p)->device);                             \
    } while (0)

// pass A on this rank's rows, then pack the blocks of exchange #1 (peer-major) into sendbuf
extern "C" int ilm_slab_forward(ilm_plan* p, const ilm_slab_info* s, int layout1, const double* w1_rows, int layout2,
                                const double* w2_rows, double* sendbuf) {
    ILM_SLAB_PLAN(p);
    if (!scalar_layout(layout1) || !w1_rows || (w2_rows && !scalar_layout(layout2)) || !sendbuf) { set_error("ilm_slab_forward: bad arguments"); return ILM_EINVAL; }
    if (!is_device_ptr(w1_rows) || !is_device_ptr(sendbuf)) { set_error("ilm_slab_forward: device-resident buffers only"); return ILM_EINVAL; }
    if (!w2_rows) layout2 = -1;
    ILM_TRY(check_info(p, s, layout_rows(p, layout1), layout_rows(p, layout2)));
    // compact geometry: rows are counted from row0, the spectrum of the nrows rows IS the peer-major send buffer
    const int nrows = s->row1 - s->row0;
    if (nrows == 0) return ILM_OK;
    ConvArgs a = conv_base_args(p);
    a.g = ConvGeom{p->Lx, p->Ly, nrows, nrows};
    a.f1 = local_ref(p, layout1, const_cast<double*>(w1_rows), s->row0, nrows);
    a.f2 = local_ref(p, layout2, const_cast<double*>(w2_rows), s->row0, nrows);
    a.rlo = 0; a.rhi = nrows; a.olo = 0; a.ohi = nrows;
    a.S = reinterpret_cast<double2*>(sendbuf);
    ILM_TRY(conv_launcher(p->Lx)(0, a, p->nsm, p->stream, nullptr));
    p->launches++;
    return ILM_OK;
}

// unpack exchange #1, pass B on this rank's columns, pack exchange #2
extern "C" int ilm_slab_columns(ilm_plan* p, const ilm_slab_info* s, int kernel_id, const double* recvbuf, double* sendbuf) {
    ILM_SLAB_PLAN(p);
    if (!s || !recvbuf || !sendbuf || !is_device_ptr(recvbuf) || !is_device_ptr(sendbuf)) { set_error("ilm_slab_columns: bad arguments"); return ILM_EINVAL; }
    if (kernel_id < 0 || kernel_id >= (int)p->kernels.size()) { set_error("ilm_slab_columns: unknown kernel id"); return ILM_EINVAL; }
    if (s->Lx != p->Lx || s->Ly != p->Ly) { set_error("ilm_slab_columns: slab info of another grid"); return ILM_EINVAL; }
    const int ntc = s->tc1 - s->tc0;
    // this rank's tile columns only; the kernels index the full layout, so the base pointers are shifted by tc0 columns
    const size_t need = (size_t)(ntc > 0 ? ntc : 1) * 2 * s->MYp;
    if (need > p->slab_cap) {
        cudaFree(p->slab_S); cudaFree(p->slab_S2);
        p->slab_S = p->slab_S2 = nullptr; p->slab_cap = 0;
        ILM_CUDA(cudaMalloc(&p->slab_S, need * sizeof(double2)));
        ILM_CUDA(cudaMalloc(&p->slab_S2, need * sizeof(double2)));
        p->slab_cap = need;
    }
    double2* S = p->slab_S - (ptrdiff_t)s->tc0 * 2 * s->MYp;
    double2* S2 = p->slab_S2 - (ptrdiff_t)s->tc0 * 2 * s->MYp;
    const double2* in = reinterpret_cast<const double2*>(recvbuf);
    for (int r = 0; r < s->nranks; ++r) {
        const Peer q = peer_of(s, p->g.NX, p->g.NY, r);
        ILM_TRY(copy_block(p, S, s->MYp, s->tc0, ntc, q.row0, q.nrows, const_cast<double2*>(in), false));
        in += (size_t)2 * ntc * q.nrows;
    }
    ConvArgs a = base_args(p, s, kernel_id);
    a.S = S; a.S2 = S2;
    a.wlo = s->wlo; a.whi = s->whi;
    if (s->whi > s->wlo) {
        ILM_TRY(conv_launcher(p->Ly)(1, a, p->nsm, p->stream, nullptr));
        p->launches++;
    }
    double2* out = reinterpret_cast<double2*>(sendbuf);
    for (int r = 0; r < s->nranks; ++r) {
        const Peer q = peer_of(s, p->g.NX, p->g.NY, r);
        ILM_TRY(copy_block(p, S2, s->MYp, s->tc0, ntc, q.row0, q.nrows, out, true));
        out += (size_t)2 * ntc * q.nrows;
    }
    return ILM_OK;
}

// unpack exchange #2, pass C on this rank's rows
extern "C" int ilm_slab_inverse(ilm_plan* p, const ilm_slab_info* s, const double* recvbuf, int layout1, double* w1_rows, int layout2,
                                double* w2_rows) {
    ILM_SLAB_PLAN(p);
    if (!scalar_layout(layout1) || !w1_rows || (w2_rows && !scalar_layout(layout2)) || !recvbuf) { set_error("ilm_slab_inverse: bad arguments"); return ILM_EINVAL; }
    if (!is_device_ptr(w1_rows) || !is_device_ptr(recvbuf)) { set_error("ilm_slab_inverse: device-resident buffers only"); return ILM_EINVAL; }
    if (!w2_rows) layout2 = -1;
    ILM_TRY(check_info(p, s, layout_rows(p, layout1), layout_rows(p, layout2)));
    // the receive buffer of exchange #2 is [tile column][my row pairs]: the compact spectrum that pass C inverts
    const int nrows = s->row1 - s->row0;
    if (nrows == 0) return ILM_OK;
    ConvArgs a = conv_base_args(p);
    a.g = ConvGeom{p->Lx, p->Ly, nrows, nrows};
    a.Ghat = p->kernels[0].ghat;
    a.f1 = local_ref(p, layout1, w1_rows, s->row0, nrows);
    a.f2 = local_ref(p, layout2, w2_rows, s->row0, nrows);
    a.rlo = 0; a.rhi = nrows; a.olo = 0; a.ohi = nrows;
    a.S2 = const_cast<double2*>(reinterpret_cast<const double2*>(recvbuf));
    if (p->Lx >= 512 && p->Lx <= 4096) ILM_TRY(make_s2_tensor_map(p, nrows, a.S2));
    ILM_TRY(conv_launcher(p->Lx)(2, a, p->nsm, p->stream, p->tmap_s2));
    p->launches++;
    return ILM_OK;
}


// the three stages with both exchanges issued inside the library (grouped ncclSend / ncclRecv on the plan's stream)
extern "C" int ilm_slab_solve(ilm_plan* p, const ilm_slab_info* s, int kernel_id, int layout1, double* w1_rows, int layout2,
                              double* w2_rows, double* sendbuf, double* recvbuf) {
    ILM_SLAB_PLAN(p);
    if (!s || !sendbuf || !recvbuf) { set_error("ilm_slab_solve: bad arguments"); return ILM_EINVAL; }
    if (!p->comm || p->comm_size != s->nranks || p->comm_rank != s->rank) {
        set_error("ilm_slab_solve: the plan's communicator does not match the slab partition (ilm_comm_init)");
        return ILM_ENCCL;
    }
    std::vector<int64_t> sc(s->nranks), rc(s->nranks);
    ILM_TRY(ilm_slab_forward(p, s, layout1, w1_rows, layout2, w2_rows, sendbuf));
    ILM_TRY(ilm_slab_counts(p->g.NX, p->g.NY, s, 0, sc.data(), rc.data()));
    ILM_TRY(comm_alltoallv(p, sendbuf, sc.data(), recvbuf, rc.data()));
    ILM_TRY(ilm_slab_columns(p, s, kernel_id, recvbuf, sendbuf));
    ILM_TRY(ilm_slab_counts(p->g.NX, p->g.NY, s, 1, sc.data(), rc.data()));
    ILM_TRY(comm_alltoallv(p, sendbuf, sc.data(), recvbuf, rc.data()));
    return ilm_slab_inverse(p, s, recvbuf, layout1, w1_rows, layout2, w2_rows);
}

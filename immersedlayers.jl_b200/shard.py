"""Multi-GPU sharding of the Schur-complement build.

The N columns of S = -E L^-1 R are independent Poisson solves
(src/matrix_operators.jl:16-26), so rank r builds a contiguous block of columns
with its own replica of the plan and the blocks are exchanged with ONE
all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no
other data-path collective."""
from __future__ import annotations


def column_ranges(n, world_size):
    """Contiguous, even-sized (so that column pairs stay on one rank) blocks."""
    pairs = (n + 1) // 2
    base, rem = divmod(pairs, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < rem else 0)
        c0 = min(2 * start, n)
        c1 = min(2 * (start + cnt), n)
        out.append((c0, c1))
        start += cnt
    return out


def allgather_columns(block, n, ranges, group=None):
    """block: torch tensor holding this rank's N x ncols column-major block as a
    flat buffer.  Returns the flat column-major N x N matrix on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    maxcols = max(c1 - c0 for c0, c1 in ranges)
    pad = torch.zeros(n * maxcols, dtype=block.dtype, device=block.device)
    pad[: block.numel()] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: n * (ranges[r][1] - ranges[r][0])] for r in range(world)])


def create_schur_sharded(builder, cache, scale=1.0, group=None):
    """builder: one of api.create_RTLinvR / create_CLinvCT / create_GLinvD.
    Every rank gets the full N x N matrix (column-major torch view)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ranges = column_ranges(cache.N, world)
    blk = builder(cache, scale=scale, cols=ranges[rank])      # N x ncols column-major view
    flat = blk.t().contiguous().reshape(-1) if blk.numel() else blk.reshape(-1)
    full = allgather_columns(flat, cache.N, ranges, group)
    return full.view(cache.N, cache.N).t()


# --------------------------------------------------------------------------
# slab decomposition of ONE convolution
# --------------------------------------------------------------------------
def slab_info(NX, NY, rows, world_size, rank):
    """Partition of one L^-1 solve over `world_size` GPUs (ilm_slab_partition; pure host
    arithmetic, no GPU needed): field rows [row0,row1) and x-frequency tile columns [tc0,tc1)."""
    from . import _lib as L
    import ctypes as C
    info = L.ilm_slab_info()
    L.check(L.load().ilm_slab_partition(NX, NY, rows, world_size, rank, C.byref(info)))
    return info


def slab_counts(NX, NY, info, phase):
    """(send_counts, recv_counts) in doubles of all-to-all #1 (phase 0) / #2 (phase 1)."""
    from . import _lib as L
    import ctypes as C
    send = (C.c_int64 * info.nranks)()
    recv = (C.c_int64 * info.nranks)()
    L.check(L.load().ilm_slab_counts(NX, NY, C.byref(info), phase, send, recv))
    return list(send), list(recv)


class SlabLaplacian:
    """inverse_laplacian! (src/grid_operators.jl:153-179) of ONE grid whose rows are spread over
    the ranks of a torch.distributed group, one process per GPU.  Every rank holds a replica of the
    plan (tables, Ghat) and a slab of rows of the field; the two transposes of the 2-D transform are
    variable-count all-to-alls (NCCL over NVLink/NVSwitch), everything else runs in libilm_b200.

        slab = SlabLaplacian(cache, L.NODES_PRIMAL)
        mine = slab.scatter(full_field)            # torch tensor with this rank's rows
        slab.inverse_laplacian(mine)               # in place
    """

    def __init__(self, cache, layout1, layout2=None, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib as L
        self.cache, self.group = cache, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.layouts = (layout1, layout2)
        g = cache.g
        shapes = [g.layout_shape(l) for l in self.layouts if l is not None]
        rows = max(s[1] for s in shapes)
        self.info = slab_info(g.NX, g.NY, rows, self.world, self.rank)
        self.counts = [slab_counts(g.NX, g.NY, self.info, ph) for ph in (0, 1)]
        n = int(L.load().ilm_slab_buffer_doubles(C.byref(self.info)))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.send = torch.empty(n, dtype=torch.float64, device=dev)
        self.recv = torch.empty(n, dtype=torch.float64, device=dev)
        self._L, self._C = L, C

    def rows(self, layout):
        """Rows [r0, r1) of a field of this layout owned by this rank."""
        my = self.cache.g.layout_shape(layout)[1]
        return min(self.info.row0, my), min(self.info.row1, my)

    def scatter(self, full, layout=None):
        """This rank's rows of a full (mx, my) array as a flat device tensor (x fastest)."""
        import numpy as np
        import torch
        layout = self.layouts[0] if layout is None else layout
        r0, r1 = self.rows(layout)
        part = np.ascontiguousarray(np.asarray(full, dtype=np.float64)[:, r0:r1].T).reshape(-1)
        return torch.from_numpy(part).to(self.send.device)

    def gather(self, mine, layout=None):
        """Full (mx, my) numpy array on every rank (test / inspection helper)."""
        import numpy as np
        import torch
        import torch.distributed as dist
        layout = self.layouts[0] if layout is None else layout
        mx, my = self.cache.g.layout_shape(layout)
        parts = [None] * self.world
        dist.all_gather_object(parts, mine.cpu().numpy(), group=self.group)
        return np.concatenate([p.reshape(-1, mx) for p in parts], axis=0).T.copy()

    def _exchange(self, phase):
        import torch.distributed as dist
        send, recv = self.counts[phase]
        dist.all_to_all_single(self.recv[: sum(recv)], self.send[: sum(send)], output_split_sizes=recv,
                               input_split_sizes=send, group=self.group)

    def inverse_laplacian(self, w1, w2=None, kernel_id=0, in_library=False):
        """w1 (and w2, the second layout) hold this rank's rows and are overwritten with L^-1 w
        (kernel_id selects another convolution kernel, e.g. an integrating factor).
        in_library=True: one call of ilm_slab_solve -- both exchanges are grouped ncclSend / ncclRecv issued by
        the library on the communicator bound to the plan (cache.comm_init); otherwise the three stages with
        torch.distributed all_to_all_single between them (the cross-check path)."""
        L, C = self._L, self._C
        lib, plan, info = self.cache._lib, self.cache._plan, C.byref(self.info)
        l1, l2 = self.layouts
        if (w2 is None) != (l2 is None):
            raise L.MethodError("SlabLaplacian: second field does not match the configured layouts")
        p1 = C.c_void_p(w1.data_ptr())
        p2 = C.c_void_p(w2.data_ptr()) if w2 is not None else None
        if in_library:
            L.check(lib.ilm_slab_solve(plan, info, kernel_id, l1, p1, -1 if l2 is None else l2, p2,
                                       C.c_void_p(self.send.data_ptr()), C.c_void_p(self.recv.data_ptr())))
            return w1 if w2 is None else (w1, w2)
        L.check(lib.ilm_slab_forward(plan, info, l1, p1, -1 if l2 is None else l2, p2, C.c_void_p(self.send.data_ptr())))
        self._exchange(0)
        L.check(lib.ilm_slab_columns(plan, info, kernel_id, C.c_void_p(self.recv.data_ptr()), C.c_void_p(self.send.data_ptr())))
        self._exchange(1)
        L.check(lib.ilm_slab_inverse(plan, info, C.c_void_p(self.recv.data_ptr()), l1, p1, -1 if l2 is None else l2, p2))
        return w1 if w2 is None else (w1, w2)


def slab_solve_virtual_ranks(cache, layouts, fields, P, kernel_id=0):
    """The slab-decomposed solve for P *virtual* ranks on ONE GPU and one plan: the three CUDA stages
    run rank after rank and the two all-to-alls are done by slicing the packed buffers exactly as
    all_to_all_single would.  Exercises partition, counts, pack / unpack and the ranged column pass
    without a multi-GPU box.  fields: full (mx, my) numpy arrays, one per layout; returns the solved
    full arrays."""
    import ctypes as C
    import numpy as np
    from . import _lib as L
    import torch
    g = cache.g
    lib, plan = cache._lib, cache._plan
    rows = max(g.layout_shape(l)[1] for l in layouts)
    infos = [slab_info(g.NX, g.NY, rows, P, r) for r in range(P)]
    counts = [[slab_counts(g.NX, g.NY, i, ph) for i in infos] for ph in (0, 1)]
    dev = torch.device("cuda")
    nbuf = max(int(lib.ilm_slab_buffer_doubles(C.byref(i))) for i in infos)
    send = [torch.zeros(nbuf, dtype=torch.float64, device=dev) for _ in range(P)]
    recv = [torch.zeros(nbuf, dtype=torch.float64, device=dev) for _ in range(P)]

    def slab(arr, layout, info):
        my = g.layout_shape(layout)[1]
        r0, r1 = min(info.row0, my), min(info.row1, my)
        return torch.from_numpy(np.ascontiguousarray(arr[:, r0:r1].T).reshape(-1)).to(dev)

    mine = [[slab(f, l, i) for f, l in zip(fields, layouts)] for i in infos]
    l1 = layouts[0]
    l2 = layouts[1] if len(layouts) > 1 else -1

    def ptr(t):
        return C.c_void_p(t.data_ptr())

    def all_to_all(phase):
        for r in range(P):
            off = 0
            for src in range(P):
                n = counts[phase][r][1][src]
                soff = sum(counts[phase][src][0][:r])
                assert counts[phase][src][0][r] == n
                recv[r][off:off + n] = send[src][soff:soff + n]
                off += n

    for r, info in enumerate(infos):
        w2 = ptr(mine[r][1]) if len(layouts) > 1 else None
        L.check(lib.ilm_slab_forward(plan, C.byref(info), l1, ptr(mine[r][0]), l2, w2, ptr(send[r])))
    all_to_all(0)
    for r, info in enumerate(infos):
        L.check(lib.ilm_slab_columns(plan, C.byref(info), kernel_id, ptr(recv[r]), ptr(send[r])))
    all_to_all(1)
    for r, info in enumerate(infos):
        w2 = ptr(mine[r][1]) if len(layouts) > 1 else None
        L.check(lib.ilm_slab_inverse(plan, C.byref(info), ptr(recv[r]), l1, ptr(mine[r][0]), l2, w2))
    out = []
    for k, l in enumerate(layouts):
        mx = g.layout_shape(l)[0]
        out.append(np.concatenate([mine[r][k].cpu().numpy().reshape(-1, mx) for r in range(P)], axis=0).T)
    return out

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

"""Cross-rank reduction of augment results with torch.distributed.

One process per GPU.  The data path has no collective; the only exchange is this
one-shot reduction after the last chunk (`reduce_results`): a reduce(SUM) to rank 0
of the int64 counter buffer [NC | IL0adj | OLadj | RC | rej, n_lines], a reduce(MIN)
of the first-touch stamps and one all_gather of the two small side tables (NCCL
over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .counts import FlatResult, merge_side

ERR_NONE = (1 << 63) - 1
ERR_HOST = 255          # error word code: a rank failed on the host side (no device error code)


def reduce_error(err_word: int, device, group=None) -> int:
    """min over ranks of (offset << 8 | code); ERR_NONE if no rank saw an error."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([err_word], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def allreduce_results(sums, stamps, novel, sparse, n_nodes: int, n_edges: int, group=None) -> FlatResult:
    """sums/stamps/novel/sparse: torch tensors (CUDA under NCCL, CPU under gloo).
    Returns the job-wide FlatResult on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(stamps, op=dist.ReduceOp.MIN, group=group)

    def gather_rows(rows):
        rows = rows.reshape(-1, 3).contiguous()
        n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(s.item()) for s in sizes]
        m = max(max(sizes), 1)
        pad = torch.zeros((m, 3), dtype=rows.dtype, device=rows.device)
        pad[: rows.shape[0]] = rows
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        parts = [b[:s].cpu().numpy().view(np.uint64).reshape(-1, 3) for b, s in zip(bufs, sizes)]
        return merge_side(parts)

    return FlatResult(n_nodes, n_edges, sums.cpu().numpy(), stamps.cpu().numpy(),
                      gather_rows(novel), gather_rows(sparse))


def gather_side(rows, group=None):
    """All ranks' {key, count, stamp} rows merged by key (counts add, stamps take the minimum), as a tensor on the rows'
    device, on every rank.  Everything stays on the device: one host read (the row counts) sizes the exchange."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rows = rows.reshape(-1, 3).contiguous()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = torch.cat(sizes).tolist()
    m = max(max(sizes), 1)
    pad = torch.zeros((m, 3), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    allr = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    if allr.shape[0] == 0:
        return allr
    uniq, inv = torch.unique(allr[:, 0], return_inverse=True)
    cnt = torch.zeros_like(uniq).scatter_add_(0, inv, allr[:, 1])
    st = torch.full_like(uniq, (1 << 63) - 1).scatter_reduce_(0, inv, allr[:, 2], "amin", include_self=True)
    return torch.stack([uniq, cnt, st], dim=1)


# first-touch stamps: below this many bytes the whole array is reduced (no dependency on the side rows, overlaps the
# counter reduction); above it only the stamps the writer can ever read are
FULL_STAMPS_BYTES = 256 << 20


def reduce_results(sums, stamps, novel, sparse, n_nodes: int, dst: int = 0, group=None):
    """The one-shot reduction that ends a multi-GPU job, to rank `dst` only (it alone writes the GFA):

    * ``reduce(SUM)`` of the counter buffer (not all_reduce: nobody else needs it), started first and asynchronously so
      that the host round trip below (the row counts size the side exchange) overlaps it;
    * the two small side tables travel in ONE all_gather; their rows are concatenated, not merged -- `rows_to_host`
      merges by key when rank `dst` takes them to the host;
    * first-touch stamps: ``reduce(MIN)`` of the whole array when it is small; for big graphs only the stamps of nodes
      that also have a deletion-derived key are reduced (the writer orders a node's IL / OL entries by them and reads
      no other stamp) -- a few thousand values instead of 2 x n_nodes.

    -> (sums, stamps, novel rows, sparse rows), tensors on the inputs' device (rows as int64 bit patterns of the uint64
    layout, one row per (rank, key)); sums / stamps are valid on `dst` only."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    novel = novel.reshape(-1, 3)
    sparse = sparse.reshape(-1, 3)
    n_mine = torch.tensor([novel.shape[0], sparse.shape[0]], dtype=torch.int64, device=sums.device)
    sizes = torch.empty(world * 2, dtype=torch.int64, device=sums.device)     # (flat buffers: gloo insists)
    dist.all_gather_into_tensor(sizes, n_mine, group=group)
    pending = [dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    full_stamps = stamps.numel() * stamps.element_size() <= FULL_STAMPS_BYTES
    if full_stamps:
        pending.append(dist.reduce(stamps, dst=dst, op=dist.ReduceOp.MIN, group=group, async_op=True))
    sizes = sizes.view(world, 2).tolist()                   # the one host read of the epilogue
    m = max(max(a + b for a, b in sizes), 1)
    pad = torch.zeros((m, 3), dtype=novel.dtype, device=sums.device)
    pad[: novel.shape[0]] = novel
    pad[novel.shape[0]: novel.shape[0] + sparse.shape[0]] = sparse
    rows = torch.empty(world * m * 3, dtype=novel.dtype, device=sums.device)
    dist.all_gather_into_tensor(rows, pad.view(-1), group=group)
    rows = rows.view(world, m, 3)
    novel_all = torch.cat([rows[r, :a] for r, (a, b) in enumerate(sizes)], dim=0)
    sparse_all = torch.cat([rows[r, a: a + b] for r, (a, b) in enumerate(sizes)], dim=0)
    if not full_stamps and sparse_all.shape[0]:
        nodes = torch.unique((sparse_all[:, 0] >> 32) & 0xFFFFFFFF)   # key = idx << 32 | ...; the same list on every rank
        idx = torch.cat([nodes, nodes + n_nodes])
        sub = stamps[idx].contiguous()
        dist.reduce(sub, dst=dst, op=dist.ReduceOp.MIN, group=group)
        if dist.get_rank(group) == dst:
            stamps[idx] = sub
    for w in pending:
        w.wait()
    return sums, stamps, novel_all, sparse_all


def rows_to_host(rows) -> np.ndarray:
    """{key, count, stamp} rows (int64 bit patterns, any device; several rows per key after `reduce_results`) ->
    uint64[n, 3] on the host, one row per key: counts add, stamps take the minimum."""
    return merge_side([rows.cpu().numpy().view(np.uint64).reshape(-1, 3)])

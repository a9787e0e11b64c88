"""Exact cross-rank merge of the per-GPU count tables -- the multi-GPU analogue of the reference's serial
map merge in process_output (src/kmer.cpp:1486-1515): element-wise sum of the six (k, seq) -> count maps.

Per-GPU open-addressing tables are not slot-aligned, so the tables cannot be reduced in place.  Recipe
(SURVEY.md 5.8): all_gather the compacted keys -> identical sorted union on every rank -> dense count
vectors aligned to the union -> reduce(SUM) to rank 0.  Integer sums are order-independent, so the result
is bit-exact.  torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests); there
is no data-path collective besides this end-of-file exchange.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def tables_to_rows(tables) -> np.ndarray:
    """{(table, k, seq): count} -> (n, 4) int64 rows (meta, seq_lo, seq_hi, count)."""
    rows = np.zeros((len(tables), 4), dtype=np.uint64)
    for i, ((tb, k, seq), c) in enumerate(sorted(tables.items())):
        rows[i] = ((tb << 8) | k, seq & (2 ** 64 - 1), seq >> 64, c)
    return rows.view(np.int64)


def rows_to_tables(rows: np.ndarray):
    r = rows.view(np.uint64)
    return {(int(m) >> 8, int(m) & 0xff, (int(hi) << 64) | int(lo)): int(c) for m, lo, hi, c in r}


def merge_rows(rows: np.ndarray, device: torch.device, dst: int = 0) -> Optional[np.ndarray]:
    """Sum the per-rank row sets by key.  Returns the merged rows on rank `dst`, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    world = dist.get_world_size()
    local = torch.from_numpy(np.ascontiguousarray(rows)).to(device)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(max(sizes), 1)
    padded = torch.zeros((n_max, 3), dtype=torch.int64, device=device)
    padded[: local.shape[0]] = local[:, :3]
    gathered = [torch.zeros((n_max, 3), dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(gathered, padded)
    keys = torch.cat([g[:s] for g, s in zip(gathered, sizes)], dim=0)
    if keys.shape[0] == 0:
        return np.zeros((0, 4), dtype=np.int64) if dist.get_rank() == dst else None
    union, inverse = torch.unique(keys, dim=0, return_inverse=True)  # sorted, identical on every rank
    start = sum(sizes[: dist.get_rank()])
    mine = inverse[start:start + local.shape[0]]
    dense = torch.zeros(union.shape[0], dtype=torch.int64, device=device)
    dense.index_add_(0, mine, local[:, 3])
    dist.reduce(dense, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    out = torch.cat([union, dense[:, None]], dim=1).cpu().numpy()
    # torch.unique sorts signed; re-sort as unsigned (table, k, seq_hi, seq_lo) like trew_dev_finish
    u = out.view(np.uint64)
    order = np.lexsort((u[:, 1], u[:, 2], u[:, 0]))
    return out[order]


def merge_device(ctx, device: torch.device, dst: int = 0) -> None:
    """Device-side exact merge for one process per GPU: every rank copies its compacted table (32-byte trew_entry rows) into a
    padded buffer, the buffers are gathered to rank `dst` over NCCL, and `dst` adds the other ranks' rows to its own
    count table with trew_dev_merge_rows (integer atomics).  Afterwards `ctx.finish*()` on rank `dst` returns the
    merged tables.  Two small collectives per file; nothing but the rows crosses NVLink."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world, rank = dist.get_world_size(), dist.get_rank()
    n = ctx.export_rows()
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=device))
    sizes = [int(s.item()) for s in sizes]
    n_max = max(max(sizes), 1)
    # torch.empty: nothing is queued on torch's stream that could race with the context's own (non-blocking) stream
    # writing the rows; the padding past n is never read (sizes travel separately)
    rows = torch.empty((n_max, 4), dtype=torch.int64, device=device)
    ctx.export_rows(rows.data_ptr(), n_max)
    gathered = [torch.empty_like(rows) for _ in range(world)] if rank == dst else None
    dist.gather(rows, gathered, dst=dst)
    if rank == dst:
        torch.cuda.synchronize(device)
        ctx.reserve(sum(sizes[r] for r in range(world) if r != dst))   # the union can be world x the local table
        for r in range(world):
            if r != dst and sizes[r]:
                ctx.merge_rows(gathered[r].data_ptr(), sizes[r])   # asynchronous on the context's stream
        ctx.sync()   # the gathered buffers must outlive the merge kernels


_exchange = {}   # id(ctx) -> [capacity in rows, row buffer with a header row, gathered buffer, rows to send next time]


def _round_rows(n: int) -> int:
    return (n + n // 8 + 1023) // 1024 * 1024


def finish_merged(ctx, device: torch.device, dst: int = 0):
    """End-of-file merge without touching the count tables: every rank copies its compacted rows into a fixed-capacity
    device buffer whose first row carries the row count, ONE NCCL all-gather moves the buffers, and `dst` forms the union
    on the device (concatenate, [report filter], radix sort, sum equal keys: trew_dev_finish_merged) and copies it to the
    host.  Only the first `send` rows of the buffers travel: the largest count any rank announced last time plus an
    eighth (the same number on every rank, since all ranks see all headers), so a file's exchange moves little more than
    its rows.  A rank whose rows do not fit (negative header) makes every rank send the whole buffer -- doubled first if
    that is too small as well -- and repeat: no extra collective in the steady state.  Returns the merged entries
    (structured numpy view, sorted by (table, k, seq)) on `dst`, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ctx.finish_view()
    world, rank = dist.get_world_size(), dist.get_rank()
    n = ctx.export_rows()
    while True:
        st = _exchange.get(id(ctx))
        if st is None:
            all_n = torch.zeros(world, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(all_n, torch.tensor([n], dtype=torch.int64, device=device))
            cap = 1024
            while cap < 2 * max(all_n.tolist()) + 1024:
                cap *= 2
            st = [cap, torch.empty((cap + 1, 4), dtype=torch.int64, device=device),
                  torch.empty((world * (cap + 1), 4), dtype=torch.int64, device=device), min(cap, _round_rows(max(all_n.tolist())))]
            _exchange[id(ctx)] = st
        cap, rows, gathered, send = st
        if n <= send:
            ctx.export_rows(rows[1:].data_ptr(), cap)
            rows[0, 0] = n
        else:
            rows[0, 0] = -n
        got = gathered[: world * (send + 1)]
        dist.all_gather_into_tensor(got, rows[: send + 1])
        heads = got.view(world, send + 1, 4)[:, 0, 0].tolist()     # one small D2H; also orders the gather before the union
        if min(heads) >= 0:
            st[3] = min(cap, _round_rows(max(heads)))
            break
        need = max(-h for h in heads if h < 0)
        if need <= cap:
            st[3] = cap
            continue
        while cap < 2 * need + 1024:
            cap *= 2
        _exchange[id(ctx)] = [cap, torch.empty((cap + 1, 4), dtype=torch.int64, device=device),
                              torch.empty((world * (cap + 1), 4), dtype=torch.int64, device=device), cap]
    if rank != dst:
        return None
    base = got.data_ptr()
    stride = (send + 1) * 32
    return ctx.finish_merged_view([(base + r * stride + 32, heads[r]) for r in range(world) if r != dst])


def forget(ctx) -> None:
    """Drop the exchange buffers kept for a context (call before closing it)."""
    _exchange.pop(id(ctx), None)

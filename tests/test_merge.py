"""The N>1 path on CPU: world_size-2 gloo run of the exact table merge (the NCCL step of bench.py / the
analogue of src/kmer.cpp:1486-1515)."""
import os
import random
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_tables(seed, n):
    rng = random.Random(seed)
    shared = random.Random(99)
    t = {}
    for _ in range(n):
        r = shared if rng.random() < 0.5 else rng  # some keys on both ranks, some on one
        k = r.choice([5, 6, 31, 32, 33, 64])
        t[(r.randrange(6), k, r.getrandbits(2 * k))] = rng.randrange(1, 2 ** 40)
    return t


def worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from trew_b200 import merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = merge.tables_to_rows(make_tables(rank, 300 if rank == 0 else 0 if rank == 2 else 500))
    merged = merge.merge_rows(rows, torch.device("cpu"))
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), merged)
    else:
        assert merged is None
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(world, tmp_path):
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    from trew_b200 import merge
    got = merge.rows_to_tables(np.load(os.path.join(str(tmp_path), "merged.npy")))
    want = {}
    for r in range(world):
        for k, v in make_tables(r, 300 if r == 0 else 0 if r == 2 else 500).items():
            want[k] = want.get(k, 0) + v
    assert got == want


def test_merge_world2(tmp_path):
    run_world(2, tmp_path)


def test_merge_world3_with_an_empty_rank(tmp_path):
    run_world(3, tmp_path)


def test_rows_roundtrip_and_order():
    from trew_b200 import merge
    t = make_tables(5, 200)
    rows = merge.tables_to_rows(t)
    assert merge.rows_to_tables(rows) == t
    assert merge.merge_rows(rows, torch.device("cpu")) is rows  # single process: identity


# ---- merge.finish_merged (the one-collective exchange bench.py uses under torchrun) with a stand-in context on gloo ------

class FakeCtx:
    """What finish_merged needs of a DeviceContext, on host memory: export_rows (count / copy to a pointer) and
    finish_merged_view (here: the lists it was handed, read back through their pointers)."""

    def __init__(self, rows):
        self.rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 4)
        self.exports = 0

    def export_rows(self, d_rows=None, capacity_rows=0):
        import ctypes
        n = self.rows.shape[0]
        if d_rows is not None:
            assert n <= capacity_rows
            self.exports += 1
            if n:
                ctypes.memmove(d_rows, self.rows.ctypes.data, n * 32)
        return n

    def finish_merged_view(self, lists):
        import ctypes
        out = [self.rows.copy()]
        for ptr, n in lists:
            a = np.empty((n, 4), dtype=np.int64)
            if n:
                ctypes.memmove(a.ctypes.data, ptr, n * 32)
            out.append(a)
        return out


def exchange_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from trew_b200 import merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    sizes = [[700, 1500], [720, 1400], [3000, 200], [90, 0], [100, 30000]]   # rows per rank, step by step
    ctx = FakeCtx(np.zeros((0, 4)))
    sent = []
    for step, per_rank in enumerate(sizes):
        n = per_rank[rank]
        ctx.rows = (np.arange(4 * n, dtype=np.int64).reshape(n, 4) + 1_000_000 * (rank + 1) + 10_000_000 * step)
        got = merge.finish_merged(ctx, dev)
        st = merge._exchange[id(ctx)]
        sent.append((st[0], st[3]))
        assert st[3] <= st[0] and st[3] >= max(per_rank)        # the next exchange sends at least what was announced
        if rank == 0:
            assert len(got) == world
            for r in range(world):
                want = np.arange(4 * per_rank[r], dtype=np.int64).reshape(per_rank[r], 4) + 1_000_000 * (r + 1) + 10_000_000 * step
                assert np.array_equal(got[r], want), (step, r)
        else:
            assert got is None
    # step 1 fits the estimate of step 0 (trimmed exchange), step 2 overflows the estimate but not the buffer, step 4 overflows the buffer
    assert sent[1][0] == sent[0][0] and sent[2][0] == sent[0][0] and sent[4][0] > sent[3][0], sent
    assert sent[0][1] < sent[0][0], sent                         # less than the whole buffer travels
    merge.forget(ctx)
    dist.destroy_process_group()
    if rank == 0:
        open(os.path.join(out_dir, "exchange_ok"), "w").write("ok")


def test_finish_merged_exchange_on_gloo(tmp_path):
    """Header row, trimmed all-gather, stale estimate (whole buffer, no regrow) and regrow -- with two ranks on gloo."""
    mp.spawn(exchange_worker, args=(2, free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "exchange_ok"))

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

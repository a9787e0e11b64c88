#!/usr/bin/env python3
"""Run under torchrun with one rank per GPU: every rank scans its share of a read set, merge.finish_merged unites the
tables over NCCL, and rank 0 checks the result against one context scanning everything (with and without the report
filter, with a capacity too small for the rows, which forces the regrow path, and with a stale row estimate).  Used by tests/test_gpu_multi.py on
boxes with at least two GPUs.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/merge_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from trew_b200 import api, merge, synth  # noqa: E402


def view_to_tables(v):
    return {(int(r["table"]), int(r["k"]), (int(r["seq_hi"]) << 64) | int(r["seq_lo"])): int(r["count"]) for r in v}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    reads = synth.adversarial_short(71, 4000) + [bytes(r) for r in synth.config_short(72, 30000, telomeric=0.02, n_rate=0.004)]
    mine = reads[rank::world]
    with api.DeviceContext(api.MODE_SHORT, 5, 32, device=local) as ctx:
        ctx.submit_reads(mine)
        for step, filt in enumerate((0, 10, 0, 10)):
            ctx.set_report_filter(filt)
            if step == 1:
                merge._exchange[id(ctx)] = [16, torch.empty((17, 4), dtype=torch.int64, device=device),
                                            torch.empty((world * 17, 4), dtype=torch.int64, device=device), 16]   # too small: regrow
            if step == 3:
                merge._exchange[id(ctx)][3] = 8   # fewer rows announced last time than there are now: whole buffer, no regrow
            got = merge.finish_merged(ctx, device)
            if step == 3:
                assert merge._exchange[id(ctx)][3] < merge._exchange[id(ctx)][0]   # back to the trimmed exchange
            if rank == 0:
                with api.DeviceContext(api.MODE_SHORT, 5, 32, device=local) as one:
                    one.submit_reads(reads)
                    one.set_report_filter(filt)
                    want = one.finish()
                assert view_to_tables(got) == want, (filt, len(got), len(want))
                assert len(want) > 100
        merge.forget(ctx)
    dist.barrier()
    if rank == 0:
        print("merge_check ok: %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

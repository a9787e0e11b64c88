"""Where the end-of-step merge spends its time: torchrun --nproc-per-node N tools/merge_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from trew_b200 import api, merge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
device = torch.device("cuda", local)
T = {}
def lap(name, t0):
    torch.cuda.synchronize(device)
    t = time.perf_counter()
    T[name] = T.get(name, 0.0) + (t - t0)
    return t

with api.DeviceContext(api.MODE_SHORT, 5, 32, device=local) as ctx:
    ctx.set_report_filter(10)
    h = ctx.synth_resident(1 + 1000 * rank, 25_000_000, 150, tel_ppm=10000, half_ppm=2000, n_ppm=1000, sub_ppm=10000)
    reps = 10
    for it in range(3 + reps):
        if it == 3:
            T.clear()
        dist.barrier()
        t = time.perf_counter()
        ctx.reset(); ctx.scan_resident(h); ctx.sync()
        t = lap("reset+scan", t)
        n = ctx.export_rows()
        t = lap("export_rows (compaction)", t)
        st = merge._exchange.get(id(ctx))
        if st is None:
            merge.finish_merged(ctx, device)
            t = time.perf_counter()
            continue
        cap, rows, gathered, send = st
        ctx.export_rows(rows[1:].data_ptr(), cap)
        t = lap("copy rows", t)
        rows[0, 0] = n
        t = lap("header write", t)
        got = gathered[: world * (send + 1)]
        dist.all_gather_into_tensor(got, rows[: send + 1])
        t = lap("all_gather (%d rows sent per rank)" % send, t)
        heads = got.view(world, send + 1, 4)[:, 0, 0].tolist()
        t = lap("heads to host", t)
        if rank == 0:
            base = got.data_ptr(); stride = (send + 1) * 32
            ctx.finish_merged_view([(base + r * stride + 32, heads[r]) for r in range(world) if r != 0])
            t = lap("union on rank 0", t)
    if rank == 0:
        for k, v in T.items():
            print("%-44s %.3f ms" % (k, v / reps * 1e3))
    merge.forget(ctx)
dist.barrier()
dist.destroy_process_group()

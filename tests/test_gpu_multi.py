"""Several GPUs in one process (trew_multi, include/trew_b200.h): the C++ fan-out of chunks over devices and the
end-of-file merge on the first device must give exactly the tables of one context scanning everything -- the
reference's N consumers + map sum (src/kmer.cpp:1271-1325, 1486-1515).  On a one-GPU box the group is built from
several contexts on the same device (same code path: round-robin deal, peer copy of the rows, device-side union)."""
import gzip
import os
import subprocess

import pytest

from trew_b200 import api, synth
from test_gpu_parity import diff_msg, run_gpu

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


def test_multi_short_equals_single():
    reads = synth.adversarial_short(51, 4000)
    want = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    for devs in device_lists():
        with api.MultiContext(api.MODE_SHORT, 5, 32, devices=devs) as m:
            assert m.device_count == len(devs)
            m.submit_reads(reads, chunk_reads=300)          # 14 chunks dealt over the devices
            got = m.finish()
            st = m.stats()
            assert st.reads == len(reads)
            assert got == want, diff_msg(got, want)
            m.reset()                                       # a second file on the same group
            m.submit_reads(reads[:1500], chunk_reads=200)
            again = m.finish()
        assert again == run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads[:1500])


def test_multi_pair_and_long_equal_single():
    r1, r2 = synth.adversarial_pairs(52, 1500, read_len=150, max_unit=32, truncate_mate2=0.0)
    want = run_gpu(api.MODE_PAIR, 5, 32, 0.5, 0.8, 150, r1, r2)
    longs = synth.adversarial_long(53, 200, min_len=150, max_len=3000, max_unit=32)
    want_long = run_gpu(api.MODE_LONG, 5, 32, 0.5, 0.8, 150, longs)
    for devs in device_lists():
        with api.MultiContext(api.MODE_PAIR, 5, 32, devices=devs) as m:
            m.submit_reads(r1, r2, chunk_reads=128)         # mates travel together
            got = m.finish()
        assert got == want, diff_msg(got, want)
        with api.MultiContext(api.MODE_LONG, 5, 32, devices=devs) as m:
            m.submit_reads(longs, chunk_reads=17)
            got = m.finish()
        assert got == want_long, diff_msg(got, want_long)


def test_multi_idle_devices_and_empty_input():
    reads = synth.adversarial_short(54, 300)
    want = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    with api.MultiContext(api.MODE_SHORT, 5, 32, devices=[0, 0, 0, 0]) as m:
        assert m.finish() == {}
        m.submit_reads(reads)                               # one chunk: three contexts stay empty
        assert m.finish() == want


def test_multi_process_file_and_cli(tmp_path):
    reads = synth.adversarial_short(55, 6000, lengths=[100, 150, 151])
    d = str(tmp_path)
    plain = os.path.join(d, "a.fastq")
    data = synth.fastq_bytes(reads)
    open(plain, "wb").write(data)
    gz = plain + ".gz"
    with gzip.open(gz, "wb") as f:
        f.write(data)
    bgz = os.path.join(d, "a.fastq.bgz")
    synth.bgzf_write(plain, bgz)
    want = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    for devs in device_lists():
        with api.MultiContext(api.MODE_SHORT, 5, 32, devices=devs, staging_bytes=1 << 18) as m:   # small slots: many batches per block
            for p in (plain, gz, bgz):
                m.reset()
                m.process_file(p)
                got = m.finish()
                assert got == want, (p, diff_msg(got, want))
    # the command line: one device, the same device three times, every visible device -- identical stdout
    outs = []
    for spec in ("0", "0,0,0", "all"):
        env = dict(os.environ, TREW_DEVICES=spec)
        outs.append(subprocess.run([api.CLI_PATH, "short", "5", "32", plain, gz], capture_output=True, check=True, env=env).stdout)
    assert outs[0] == outs[1] == outs[2] and outs[0].count(b"\n") > 6


def test_nccl_merge_across_ranks():
    """One process per GPU: merge.finish_merged over NCCL gives rank 0 the tables of one context scanning everything
    (tools/merge_check.py under torchrun; needs at least two GPUs, so the single-GPU box skips it)."""
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "merge_check.py")],
                       capture_output=True, timeout=600)
    assert r.returncode == 0 and b"merge_check ok" in r.stdout, r.stderr.decode()[-2000:]

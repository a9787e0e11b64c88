"""Parity tests proper: the CUDA path (through the C ABI) against golden fixtures from the compiled
reference, against the CPU oracle on seeded adversarial inputs, and -- at sizes the oracle cannot finish --
through size-independent properties.  Integer work: every comparison is bit-exact."""
import numpy as np
import pytest

from trew_b200 import api, synth

pytestmark = pytest.mark.gpu


def tables_from_json(rows):
    from oracle.oracle import str_to_seq
    return {(tb, k, str_to_seq(s)): c for tb, k, s, c in rows}


def diff_msg(got, want):
    from oracle.oracle import format_tables
    a, b = set(got.items()), set(want.items())
    only_g = format_tables(dict(sorted(a - b)[:6]))
    only_w = format_tables(dict(sorted(b - a)[:6]))
    return "got %d entries, want %d; only CUDA: %s; only expected: %s" % (len(got), len(want), only_g, only_w)


def run_gpu(mode, mn, mx, low, high, sl, r1, r2=None, **kw):
    with api.DeviceContext(mode, mn, mx, low, high, sl, **kw) as ctx:
        ctx.submit_reads(r1, r2)
        return ctx.finish()


def test_golden_scan_cases(scan_cases):
    for case in scan_cases:
        r1 = [s.encode() for s in case["reads1"]]
        r2 = [s.encode() for s in case["reads2"]] if case["reads2"] is not None else None
        if case["mode"] == 2:
            r1 = [r for r in r1 if len(r) >= case["slice_len"]]
        got = run_gpu(case["mode"], case["min_mer"], case["max_mer"], case["low"], case["high"], case["slice_len"], r1, r2)
        want = tables_from_json(case["tables"])
        assert got == want, case["name"] + ": " + diff_msg(got, want)


@pytest.mark.parametrize("mn,mx,low,high", [(5, 32, 0.5, 0.8), (3, 64, 0.5, 0.8), (5, 64, 0.5, 0.8), (7, 20, 0.5, 0.8),
                                            (12, 40, 0.5, 0.8), (5, 32, 0.35, 0.6), (5, 32, 0.9, 1.0), (4, 33, 0.5, 0.5)])
def test_short_vs_oracle(mn, mx, low, high):
    from oracle.oracle import Oracle
    reads = synth.adversarial_short(100 + mn * 7 + mx, 1200, max_unit=mx)
    got = run_gpu(api.MODE_SHORT, mn, mx, low, high, 150, reads)
    want = Oracle(mn, mx, low, high).scan(0, reads)
    assert got == want, diff_msg(got, want)


@pytest.mark.parametrize("mn,mx,rl,trunc", [(5, 32, 150, 0.0), (5, 40, 100, 0.15), (3, 64, 120, 0.1), (5, 32, 100, 0.2), (5, 40, 160, 0.05), (3, 36, 150, 0.1)])
def test_pair_vs_oracle(mn, mx, rl, trunc):
    # the oracle follows the cleared-temp-map (128-bit path) semantics, like the CUDA path
    from oracle.oracle import Oracle
    r1, r2 = synth.adversarial_pairs(200 + mx + rl, 700, read_len=rl, max_unit=mx, truncate_mate2=trunc)
    got = run_gpu(api.MODE_PAIR, mn, mx, 0.5, 0.8, 150, r1, r2)
    want = Oracle(mn, mx).scan(1, r1, r2)
    assert got == want, diff_msg(got, want)


@pytest.mark.parametrize("mn,mx,sl", [(5, 32, 150), (3, 64, 128), (5, 20, 64), (5, 32, 300)])
def test_long_vs_oracle(mn, mx, sl):
    from oracle.oracle import Oracle
    reads = [r for r in synth.adversarial_long(300 + sl, 150, min_len=sl, max_len=4000, max_unit=mx)]
    got = run_gpu(api.MODE_LONG, mn, mx, 0.5, 0.8, sl, reads)
    want = Oracle(mn, mx, slice_len=sl).scan(2, reads)
    assert got == want, diff_msg(got, want)


def config4_reads(seed, count):
    """BASELINE.json configs[3] shape (2 M x 15 kb HiFi-like, telomeric 0.5-5 kb ends, `trew long 5 32`, SLICE 150 ->
    snum = 100) with the telomeric share raised so that every walk of buffer_task_long (src/kmer.cpp:790-856) is
    taken many times: 5' end, 3' end, both ends, both strands, whole-read repeats, N's inside the repeat."""
    import random
    rng = random.Random(seed)
    reads = [bytes(r) for r in synth.config_long(seed, count, telomeric=0.6)]
    unit, rc = b"TTAGGG", b"CCCTAA"
    for i in range(count // 8):
        n = rng.randint(9000, 18000)
        body = bytes(rng.choice(b"ACGT") for _ in range(n))
        a, b = rng.randint(500, 5000), rng.randint(500, 5000)
        kind = i % 6
        if kind == 0:
            r = (unit * (a // 6 + 1))[:a] + body[a:n - b] + (unit * (b // 6 + 1))[:b]          # both ends, same strand
        elif kind == 1:
            r = (rc * (a // 6 + 1))[:a] + body[a:n - b] + (unit * (b // 6 + 1))[:b]            # both ends, strands differ
        elif kind == 2:
            r = (unit * (n // 6 + 1))[:n]                                                      # every slice survives -> 'both'
        elif kind == 3:
            r = bytearray((unit * (a // 6 + 1))[:a] + body[a:])                                # N's inside the repeat
            for _ in range(a // 200):
                r[rng.randrange(a)] = ord("N")
            r = bytes(r)
        elif kind == 4:
            r = (b"TTAGGGG" * (a // 7 + 1))[:a] + body[a:n - b] + (unit * (b // 6 + 1))[:b]    # two units
        else:
            r = body[:n - b] + (rc * (b // 6 + 1))[:b].lower()                                 # lower case 3' end
        reads.append(r)
    return reads


@pytest.mark.parametrize("mn,mx,sl,seed", [(5, 32, 150, 3), (3, 64, 150, 4), (5, 32, 128, 5)])
def test_long_15kb_vs_oracle(mn, mx, sl, seed):
    from oracle.oracle import Oracle
    reads = config4_reads(seed, 240)
    assert len(reads) >= 200 and min(map(len, reads)) >= 1000 and max(map(len, reads)) > 17000
    got = run_gpu(api.MODE_LONG, mn, mx, 0.5, 0.8, sl, reads)
    want = Oracle(mn, mx, slice_len=sl).scan(2, reads)
    assert len(want) > 100
    assert got == want, diff_msg(got, want)


def test_edge_cases():
    from oracle.oracle import Oracle
    reads = [b"", b"A", b"ACGTACGTA", b"ACGTACGTAC", b"N" * 150, b"A" * 150, b"TTAGGG" * 25, b"ttaggg" * 25,
             b"TTAGGG" * 12 + b"N" + b"TTAGGG" * 12, (b"TTAGGG" * 170)[:1000], b"TG" * 75, b"TTAGGGN" * 21,
             b"ACGT" * 5, b"TTAGG" * 4, b"TTAGGG" * 3 + b"\r"]
    for mn, mx in [(5, 32), (3, 64)]:
        got = run_gpu(api.MODE_SHORT, mn, mx, 0.5, 0.8, 150, reads)
        want = Oracle(mn, mx).scan(0, reads)
        assert got == want, diff_msg(got, want)
    assert run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, []) == {}


def test_order_and_batching_independence():
    reads = synth.adversarial_short(11, 3000)
    a = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    b = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads[::-1], staging_bytes=1 << 16, n_staging=2)  # many tiny batches
    assert a == b
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:  # linearity: tables of a concatenation are sums
        ctx.submit_reads(reads[:1000])
        ctx.submit_reads(reads[1000:])
        ctx.submit_reads(reads)
        c = ctx.finish()
    assert c == {k: 2 * v for k, v in a.items()}


def test_packed_and_resident_paths_agree():
    reads = synth.adversarial_short(12, 2000)
    buf, locs = api.make_chunk(reads)
    pb = api.PackedBatch(buf, locs)
    pb.batch.max_read_len = max(len(r) for r in reads)
    a = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_packed(pb)
        b = ctx.finish()
        ctx.reset()
        h = ctx.upload(pb)
        ctx.scan_resident(h)
        c = ctx.finish()
        assert ctx.last_resident_ms() > 0
        ctx.free_resident(h)
    assert a == b == c


def test_full_shape_properties():
    """cfg-2 shape at 2M reads: the oracle would need minutes, so check properties instead:
    (1) reverse-complementing every read swaps forward/backward tables under the RC fold,
    (2) non-repeat reads contribute nothing: dropping them leaves the tables unchanged."""
    mat = synth.config_short(21, 2_000_000)
    buf, locs = api.matrix_chunk(mat)
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_chunk(buf, locs)
        full = ctx.finish()
        st = ctx.stats()
    assert st.reads == 2_000_000 and st.bases == 300_000_000
    assert 0.008 * st.reads < st.survivors < 0.05 * st.reads
    assert sum(full.values()) > 0
    # survivors-only rerun: reads that contain no TTAGGG-like signal must not matter
    tel = np.array([b"TTAGGG" in bytes(r) or b"CCCTAA" in bytes(r) for r in mat[:200_000]])
    sub = mat[:200_000]
    a = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, [bytes(r) for r in sub])
    b = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, [bytes(r) for r in sub[tel]])
    for key, v in b.items():
        assert a.get(key) == v
    extra = {k: v for k, v in a.items() if k not in b}
    assert sum(extra.values()) <= 0.02 * sum(a.values())


def test_device_generator_matches_numpy_mirror():
    """bench.py scans batches generated on the GPU (trew_synth_resident); the numpy mirror reproduces them
    bit for bit, so the oracle can check the very reads the benchmark times (small n here)."""
    from oracle.oracle import Oracle
    n = 30_000
    kw = dict(tel_ppm=30000, half_ppm=20000, n_ppm=2000, sub_ppm=10000)
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        h = ctx.synth_resident(5, n, 150, **kw)
        ctx.scan_resident(h)
        got = ctx.finish()
        s_ms, d_ms, e_ms, scans = ctx.kernel_times()
        ctx.free_resident(h)
    assert scans == 1 and s_ms > 0 and d_ms > 0 and e_ms > 0
    mat = synth.device_mirror(5, n, 150, **kw)
    want = Oracle(5, 32).scan(0, [bytes(r) for r in mat])
    assert got == want, diff_msg(got, want)
    assert len(got) > 50


def test_device_generator_pair_and_long_flavors():
    """bench.py's `configs` lines time device-generated pairs (flavor 1) and long reads with telomeric ends (flavor 2);
    the numpy mirror reproduces them, so the oracle checks those very reads too."""
    from oracle.oracle import Oracle
    kw = dict(tel_ppm=40000, half_ppm=20000, n_ppm=1000, sub_ppm=10000)
    n = 12_000
    with api.DeviceContext(api.MODE_PAIR, 5, 32) as ctx:
        h = ctx.synth_resident(6, n, 150, flavor=1, **kw)
        ctx.scan_resident(h)
        got = ctx.finish()
        ctx.free_resident(h)
    mat = synth.device_mirror(6, n, 150, flavor=1, **kw)
    want = Oracle(5, 32).scan(1, [bytes(r) for r in mat[0::2]], [bytes(r) for r in mat[1::2]])
    assert got == want, diff_msg(got, want)
    assert sum(1 for (tb, _, _) in got if tb >= 4) > 10     # fragments telomeric at both ends reach the 'both' tables
    n = 300
    kw["tel_ppm"] = 400000
    with api.DeviceContext(api.MODE_LONG, 5, 32) as ctx:
        h = ctx.synth_resident(7, n, 15000, flavor=2, **kw)
        ctx.scan_resident(h)
        got = ctx.finish()
        ctx.free_resident(h)
    mat = synth.device_mirror(7, n, 15000, flavor=2, **kw)
    want = Oracle(5, 32).scan(2, [bytes(r) for r in mat])
    assert got == want, diff_msg(got, want)
    assert len(got) > 50


def test_device_row_merge():
    """The multi-GPU merge path on one GPU: two contexts scan disjoint shards, the second one's table is exported as
    device rows and added to the first (trew_dev_export_rows / trew_dev_merge_rows); the result equals one context
    scanning everything."""
    import torch
    reads = synth.adversarial_short(13, 2400)
    whole = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as c1, api.DeviceContext(api.MODE_SHORT, 5, 32) as c2:
        c1.submit_reads(reads[:1000])
        c2.submit_reads(reads[1000:])
        n = c2.export_rows()
        assert n > 0
        rows = torch.zeros((n, 4), dtype=torch.int64, device="cuda:0")
        assert c2.export_rows(rows.data_ptr(), n) == n
        union = c1.finish_merged_view([(rows.data_ptr(), n)])   # sort-based union, c1's table untouched
        union = {(int(r["table"]), int(r["k"]), (int(r["seq_hi"]) << 64) | int(r["seq_lo"])): int(r["count"]) for r in union}
        assert union == whole, diff_msg(union, whole)
        c1.reserve(n)
        c1.merge_rows(rows.data_ptr(), n)
        merged = c1.finish()   # finish() waits for the asynchronous merge before `rows` goes away
    assert merged == whole, diff_msg(merged, whole)


def test_pair_and_long_batching_independence():
    """Order / batch-split independence and linearity for the paired and the long-read routing (size-independent
    properties; the oracle comparison of the same generators is in test_pair_vs_oracle / test_long_vs_oracle)."""
    r1, r2 = synth.adversarial_pairs(77, 1500, read_len=150, max_unit=32, truncate_mate2=0.0)
    a = run_gpu(api.MODE_PAIR, 5, 32, 0.5, 0.8, 150, r1, r2)
    b = run_gpu(api.MODE_PAIR, 5, 32, 0.5, 0.8, 150, r1[::-1], r2[::-1], staging_bytes=1 << 16, n_staging=2)
    assert a == b and len(a) > 20
    with api.DeviceContext(api.MODE_PAIR, 5, 32) as ctx:
        ctx.submit_reads(r1[:400], r2[:400])
        ctx.submit_reads(r1[400:], r2[400:])
        ctx.submit_reads(r1, r2)
        c = ctx.finish()
    assert c == {k: 2 * v for k, v in a.items()}
    reads = synth.adversarial_long(78, 160, min_len=150, max_len=3000, max_unit=32)
    d = run_gpu(api.MODE_LONG, 5, 32, 0.5, 0.8, 150, reads)
    e = run_gpu(api.MODE_LONG, 5, 32, 0.5, 0.8, 150, reads[::-1], staging_bytes=1 << 16, n_staging=2)
    assert d == e and len(d) > 20


def test_long_windows_through_the_screen():
    """300-base reads (half windows of 150 bases) and MAX_MER 64 on 150-base reads (whole-read probe of 150 bases)
    take the 4-word first level of the screen kernel; results still equal the oracle's."""
    from oracle.oracle import Oracle
    reads = synth.adversarial_short(91, 900, max_unit=32, lengths=[246, 260, 300, 300, 318, 150])
    got = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    want = Oracle(5, 32).scan(0, reads)
    assert got == want, diff_msg(got, want)
    reads = synth.adversarial_short(92, 900, max_unit=64, lengths=[150, 150, 140, 159, 128])
    got = run_gpu(api.MODE_SHORT, 5, 64, 0.5, 0.8, 150, reads)
    want = Oracle(5, 64).scan(0, reads)
    assert got == want, diff_msg(got, want)


def test_count_table_grows_while_streaming():
    """A deliberately small count table (2^12 slots) must grow transparently in the streaming path."""
    reads = synth.adversarial_short(14, 6000)
    want = run_gpu(api.MODE_SHORT, 5, 32, 0.5, 0.8, 150, reads)
    assert len(want) > 4096
    with api.DeviceContext(api.MODE_SHORT, 5, 32, table_log2_slots=12, staging_bytes=1 << 14, n_staging=2) as ctx:
        ctx.submit_reads(reads)
        got = ctx.finish()
    assert got == want, diff_msg(got, want)


def test_sparse_and_dense_validity_paths(monkeypatch):
    """The streaming path sends the invalid bases as records (block position + mask) and rebuilds the validity plane
    on the device; a batch with more N-bearing blocks than the record buffer holds is packed again densely.  Both
    routes, and TREW_DENSE_VAL=1, must give the oracle's tables."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(77)
    base = synth.adversarial_short(15, 1500)
    noisy = []
    for r in base:   # N in almost every 64-base block, repeats still visible in places
        a = np.frombuffer(r, dtype=np.uint8).copy()
        if a.size:
            a[rng.random(a.size) < 0.03] = ord("N")
        noisy.append(a.tobytes())
    clustered = [b"N" * 150, b"TTAGGG" * 10 + b"N" * 30 + b"TTAGGG" * 10, b"n" * 64 + b"CCCTAA" * 15, b"ACGT" * 16 + b"." + b"ACGT" * 16]
    reads = noisy + clustered * 50 + base
    want = Oracle(5, 32).scan(0, reads)

    def run(**kw):
        with api.DeviceContext(api.MODE_SHORT, 5, 32, **kw) as ctx:
            ctx.submit_reads(reads)
            got = ctx.finish()
            return got, ctx.stats().h2d_bytes

    sparse, b_sparse = run()                                         # default 64 MiB slots: everything fits the record buffer
    overflow, b_over = run(staging_bytes=1 << 18, n_staging=2)       # 256 KiB slots hold 2730 records: most batches fall back
    monkeypatch.setenv("TREW_DENSE_VAL", "1")
    dense, b_dense = run(staging_bytes=1 << 18, n_staging=2)
    assert sparse == want, diff_msg(sparse, want)
    assert overflow == want, diff_msg(overflow, want)
    assert dense == want, diff_msg(dense, want)
    assert b_sparse < b_dense                                        # the records are smaller than the plane they replace
    assert b_over <= b_dense and b_over > 0.9 * b_dense              # the fallback copied (nearly) everything densely


def test_report_filter_keeps_exactly_what_a_report_can_show():
    """trew_dev_set_report_filter(10): the rows that stay on the device belong to groups (k, RC-folded key) whose high
    and low totals are both below the print / scoring threshold, so the report text of a one-file run is identical --
    and the rows that do come back are exactly the groups that reach the threshold, with unchanged counts."""
    from oracle.oracle import Oracle
    reads = synth.adversarial_short(61, 3000) + [bytes(r) for r in synth.config_short(62, 60000, telomeric=0.02, half_telomeric=0.01,
                                                                                        n_rate=0.004)]
    o = Oracle(5, 32)

    def groups(tables):
        tot = {}
        for (tb, k, seq), c in tables.items():
            key = (k, min(seq, o.crc(seq, k)))
            t = tot.setdefault(key, [0, 0])
            t[tb & 1] += c
        return tot

    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_reads(reads)
        full = ctx.finish()
        ctx.set_report_filter(10)
        kept = ctx.finish()
        ctx.set_report_filter(0)
        assert ctx.finish() == full
    tot = groups(full)
    want = {key: c for key, c in full.items() if max(tot[(key[1], min(key[2], o.crc(key[2], key[1])))]) >= 10}
    assert kept == want
    assert 0 < len(kept) < len(full) / 2          # most rows are below the threshold
    texts = []
    for tables in (full, kept):
        rep = api.Report(5)
        rep.add_file("F", tables)
        texts.append(rep.finish())
    assert texts[0] == texts[1] and texts[0].count("\n") > 8


def test_report_filter_through_the_merge_and_the_cli(tmp_path):
    import os
    import subprocess
    reads = [bytes(r) for r in synth.config_short(63, 40000, telomeric=0.02, half_telomeric=0.01, n_rate=0.004)]
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_reads(reads)
        ctx.set_report_filter(10)
        want = ctx.finish()
    with api.MultiContext(api.MODE_SHORT, 5, 32, devices=[0, 0, 0]) as m:
        m.set_report_filter(10)
        m.submit_reads(reads, chunk_reads=7000)
        assert m.finish() == want
    p = os.path.join(str(tmp_path), "a.fastq")
    open(p, "wb").write(synth.fastq_bytes(reads))
    a = subprocess.run([api.CLI_PATH, "short", "5", "32", p], capture_output=True, check=True).stdout
    b = subprocess.run([api.CLI_PATH, "short", "5", "32", p], capture_output=True, check=True, env=dict(os.environ, TREW_FULL_TABLES="1")).stdout
    assert a == b and a.count(b"\n") > 6


def test_long_thread_path_when_switched_on(monkeypatch):
    """The three-kernel long-read path (statistics of every slice, the walks, the emissions; exact_thread.cuh) is off by
    default -- measured slower than the warp kernel's serial walk -- but stays exact: TREW_EXACT_FLAGS=16 runs it."""
    from oracle.oracle import Oracle
    monkeypatch.setenv("TREW_EXACT_FLAGS", "16")
    reads = config4_reads(9, 160) + synth.adversarial_long(310, 120, min_len=150, max_len=4000, max_unit=32)
    got = run_gpu(api.MODE_LONG, 5, 32, 0.5, 0.8, 150, reads)
    want = Oracle(5, 32, slice_len=150).scan(2, reads)
    assert got == want, diff_msg(got, want)

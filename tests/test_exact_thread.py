"""The thread-per-read exact routing (trew_b200/csrc/exact_thread.cuh) is plain scalar code that also compiles for the
host: tests/native/exact_thread_check.cpp packs reads into the device's planar layout and runs the very functions the
CUDA kernel runs.  Here its tables are diffed against the CPU oracle -- bit-exact, no GPU needed.  Reads outside the
path's limits (length, class-list capacity) must be reported as bailed and contribute nothing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Oracle
from trew_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def etc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("etc") / "libetc.so")
    src = os.path.join(ROOT, "tests", "native", "exact_thread_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    lib = C.CDLL(so)
    lib.etc_scan_reads.restype = C.c_long
    lib.etc_scan_reads.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
    lib.etc_scan_long.restype = C.c_long
    lib.etc_scan_long.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
    lib.etc_scan_pairs.restype = C.c_long
    lib.etc_scan_pairs.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
    lib.etc_scan_reads_wide.restype = C.c_long
    lib.etc_scan_reads_wide.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
    lib.etc_scan_pairs_wide.restype = C.c_long
    lib.etc_scan_pairs_wide.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
    return lib


def thr_table(B):
    """min M with (double)M / (double)T >= B, the reference's own IEEE division (src/kmer.cpp:2223-2224)."""
    t = np.full(1025, 0xFFFF, dtype=np.uint16)
    for T in range(1, 1025):
        M = np.arange(0, T + 1, dtype=np.float64)
        ok = M / np.float64(T) >= np.float64(B)
        t[T] = int(np.argmax(ok))
    return t


def run(lib, reads, mn, mx, low=0.5, high=0.8, wide=False):
    """wide: the 128-bit-key instantiation (MAX_MER <= 64); keys come back as hi << 64 | lo like the oracle's."""
    buf, locs = api.make_chunk(reads)
    locs = np.ascontiguousarray(locs, dtype=np.int32)
    tl, th = thr_table(low), thr_table(high)
    cap = 1 << 20
    ot, ok = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    okey, oc = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    nb = C.c_long(0)
    bi = np.zeros(max(1, len(reads)), np.int32)
    ohi = np.zeros(cap if wide else 1, np.uint64)
    if wide:
        n = lib.etc_scan_reads_wide(buf.tobytes(), locs.ctypes.data, len(reads), mn, mx, tl.ctypes.data, th.ctypes.data, ot.ctypes.data,
                                    ok.ctypes.data, okey.ctypes.data, ohi.ctypes.data, oc.ctypes.data, cap, C.byref(nb), bi.ctypes.data)
    else:
        n = lib.etc_scan_reads(buf.tobytes(), locs.ctypes.data, len(reads), mn, mx, tl.ctypes.data, th.ctypes.data, ot.ctypes.data,
                               ok.ctypes.data, okey.ctypes.data, oc.ctypes.data, cap, C.byref(nb), bi.ctypes.data)
    assert n >= 0
    tables = {(int(ot[i]), int(ok[i]), int(okey[i]) | ((int(ohi[i]) << 64) if wide else 0)): int(oc[i]) for i in range(n)}
    return tables, [int(x) for x in bi[:nb.value]]


@pytest.mark.parametrize("mn,mx,low,high,lengths", [(5, 32, 0.5, 0.8, [150, 150, 151, 128, 160, 149]),
                                                    (5, 25, 0.5, 0.8, [100, 101, 125, 150]),
                                                    (3, 20, 0.5, 0.8, [80, 99, 150]),
                                                    (7, 20, 0.35, 0.6, [150, 97]),
                                                    (5, 32, 0.9, 1.0, [150, 131]),
                                                    (4, 31, 0.5, 0.5, [124, 150])])
def test_thread_path_equals_oracle(etc, mn, mx, low, high, lengths):
    reads = synth.adversarial_short(900 + mn * 31 + mx, 2500, max_unit=mx, lengths=lengths)
    got, bailed = run(etc, reads, mn, mx, low, high)
    keep = [r for i, r in enumerate(reads) if i not in set(bailed)]
    want = Oracle(mn, mx, low, high).scan(0, keep)
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100
    assert bailed == [i for i, r in enumerate(reads) if len(r) > 160]   # only the length limit


@pytest.mark.parametrize("mn,mx,lengths", [(3, 64, [150, 150, 151, 160, 131]), (5, 40, [150, 100, 128]), (12, 64, [160, 150, 90]),
                                           (5, 33, [150, 140]), (5, 32, [150, 99])])
def test_wide_thread_path_equals_oracle(etc, mn, mx, lengths):
    """Units above 32 bases (128-bit keys, the reference's k_mer_check_128 path, src/kmer.cpp:2264-2328) -- and the
    same instantiation on a configuration the 64-bit one handles too."""
    reads = synth.adversarial_short(1300 + mn * 67 + mx, 2500, max_unit=mx, lengths=lengths)
    got, bailed = run(etc, reads, mn, mx, wide=True)
    assert bailed == []
    want = Oracle(mn, mx).scan(0, reads)
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100
    if mx > 32:
        assert any(k > 32 for (_, k, _) in got)
        narrow, nb = run(etc, reads, mn, mx)          # the 64-bit instantiation refuses MAX_MER > 32
        assert narrow == {} and len(nb) == len([r for r in reads if len(r) >= 2 * mn])


def test_thread_path_on_the_bench_distribution(etc):
    mat = synth.config_short(77, 30000, telomeric=0.05, half_telomeric=0.02, n_rate=0.003)
    reads = [bytes(r) for r in mat]
    got, bailed = run(etc, reads, 5, 32)
    keep = [r for i, r in enumerate(reads) if i not in set(bailed)]
    assert got == Oracle(5, 32).scan(0, keep)
    assert bailed == []


def test_thread_path_limits_and_short_reads(etc):
    # longer than 160 bases: bail; shorter than 4 * MAX_MER: the whole-read scan for the large periods runs here too
    reads = [b"TTAGGG" * 20, b"TTAGGG" * 30, b"ACGT" * 2, b"TTAGGG" * 25, (b"TTAGGGATCGATCGGCTAGCTAGGACT" * 5)[:100], b"ACGTTGCA" * 2,
             b"TTAGGG" * 3, (b"GATTACAGATTACCGATTAC" * 4)[:70]]
    got, bailed = run(etc, reads, 5, 32)
    assert bailed == [1]
    assert got == Oracle(5, 32).scan(0, [r for i, r in enumerate(reads) if i != 1]) and len(got) > 3


@pytest.mark.parametrize("mn,mx,lengths", [(5, 32, [100, 100, 75, 50, 36, 20, 12, 9]), (3, 30, [60, 90, 119, 121, 11, 6]), (7, 24, [95, 97, 40, 28, 27])])
def test_thread_path_short_reads_equal_oracle(etc, mn, mx, lengths):
    """Reads below 4 * MAX_MER: the large-k whole-read scan (src/kmer.cpp:165-171) and every guard (n < 2 MIN, n < 4 MIN)."""
    reads = synth.adversarial_short(700 + mn + mx, 3000, max_unit=mx, lengths=lengths)
    got, bailed = run(etc, reads, mn, mx)
    assert bailed == []
    want = Oracle(mn, mx).scan(0, reads)
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100


def run_pairs(lib, r1, r2, mn, mx, low=0.5, high=0.8, wide=False):
    b1, l1 = api.make_chunk(r1)
    b2, l2 = api.make_chunk(r2)
    l1 = np.ascontiguousarray(l1, dtype=np.int32)
    l2 = np.ascontiguousarray(l2, dtype=np.int32)
    tl, th = thr_table(low), thr_table(high)
    cap = 1 << 20
    ot, ok = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    okey, oc = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    nb = C.c_long(0)
    bi = np.zeros(max(1, len(r1)), np.int32)
    ohi = np.zeros(cap if wide else 1, np.uint64)
    if wide:
        n = lib.etc_scan_pairs_wide(b1.tobytes(), l1.ctypes.data, b2.tobytes(), l2.ctypes.data, len(r1), mn, mx, tl.ctypes.data,
                                    th.ctypes.data, ot.ctypes.data, ok.ctypes.data, okey.ctypes.data, ohi.ctypes.data, oc.ctypes.data, cap,
                                    C.byref(nb), bi.ctypes.data)
    else:
        n = lib.etc_scan_pairs(b1.tobytes(), l1.ctypes.data, b2.tobytes(), l2.ctypes.data, len(r1), mn, mx, tl.ctypes.data, th.ctypes.data,
                               ot.ctypes.data, ok.ctypes.data, okey.ctypes.data, oc.ctypes.data, cap, C.byref(nb), bi.ctypes.data)
    assert n >= 0
    return ({(int(ot[i]), int(ok[i]), int(okey[i]) | ((int(ohi[i]) << 64) if wide else 0)): int(oc[i]) for i in range(n)},
            [int(x) for x in bi[:nb.value]])


@pytest.mark.parametrize("mn,mx,rl,trunc", [(5, 32, 150, 0.0), (5, 32, 150, 0.15), (5, 25, 120, 0.1), (3, 20, 100, 0.2), (7, 30, 160, 0.05)])
def test_thread_pair_path_equals_oracle(etc, mn, mx, rl, trunc):
    r1, r2 = synth.adversarial_pairs(400 + mx + rl, 1500, read_len=rl, max_unit=mx, truncate_mate2=trunc)
    got, bailed = run_pairs(etc, r1, r2, mn, mx)
    skip = set(bailed)
    k1 = [r for i, r in enumerate(r1) if i not in skip]
    k2 = [r for i, r in enumerate(r2) if i not in skip]
    want = Oracle(mn, mx).scan(1, k1, k2)
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100
    # only the limits bail: a mate longer than 160 bases, or the shorter mate below 4 * MAX_MER (and at least 2 * MIN_MER)
    expect = [i for i, (a, b) in enumerate(zip(r1, r2))
              if min(len(a), len(b)) >= 2 * mn and (max(len(a), len(b)) > 160 or min(len(a), len(b)) < 4 * mx)]
    assert bailed == expect


@pytest.mark.parametrize("mn,mx,rl,trunc", [(5, 40, 160, 0.0), (3, 36, 150, 0.1), (5, 32, 150, 0.1)])
def test_wide_thread_pair_path_equals_oracle(etc, mn, mx, rl, trunc):
    r1, r2 = synth.adversarial_pairs(900 + mx + rl, 1500, read_len=rl, max_unit=mx, truncate_mate2=trunc)
    got, bailed = run_pairs(etc, r1, r2, mn, mx, wide=True)
    skip = set(bailed)
    want = Oracle(mn, mx).scan(1, [r for i, r in enumerate(r1) if i not in skip], [r for i, r in enumerate(r2) if i not in skip])
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100
    expect = [i for i, (a, b) in enumerate(zip(r1, r2))
              if min(len(a), len(b)) >= 2 * mn and (max(len(a), len(b)) > 160 or min(len(a), len(b)) < 4 * mx)]
    assert bailed == expect


def run_long(lib, reads, mn, mx, sl, low=0.5, high=0.8):
    buf, locs = api.make_chunk(reads)
    locs = np.ascontiguousarray(locs, dtype=np.int32)
    tl, th = thr_table(low), thr_table(high)
    cap = 1 << 20
    ot, ok = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    okey, oc = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    nb = C.c_long(0)
    bi = np.zeros(max(1, len(reads)), np.int32)
    n = lib.etc_scan_long(buf.tobytes(), locs.ctypes.data, len(reads), mn, mx, sl, tl.ctypes.data, th.ctypes.data, ot.ctypes.data,
                          ok.ctypes.data, okey.ctypes.data, oc.ctypes.data, cap, C.byref(nb), bi.ctypes.data)
    assert n >= 0
    return {(int(ot[i]), int(ok[i]), int(okey[i])): int(oc[i]) for i in range(n)}, [int(x) for x in bi[:nb.value]]


@pytest.mark.parametrize("mn,mx,sl,seed", [(5, 32, 150, 3), (5, 32, 128, 5), (5, 20, 64, 6), (7, 30, 160, 7)])
def test_thread_long_path_equals_oracle(etc, mn, mx, sl, seed):
    """The three-step long-read path (all slice statistics, the walks over them, the emissions) against the oracle on
    adversarial long reads and on config-4 shaped reads; reads whose walk reaches the (longer) middle slice bail out."""
    from test_gpu_parity import config4_reads
    reads = [r for r in synth.adversarial_long(500 + sl, 150, min_len=sl, max_len=4000, max_unit=mx)] + config4_reads(seed, 64)
    got, bailed = run_long(etc, reads, mn, mx, sl)
    keep = [r for i, r in enumerate(reads) if i not in set(bailed)]
    want = Oracle(mn, mx, slice_len=sl).scan(2, keep)
    assert got == want, (len(got), len(want), sorted(set(got.items()) ^ set(want.items()))[:6])
    assert len(got) > 100 and len(bailed) < len(reads) / 2


def test_thread_path_under_sanitizers(tmp_path):
    """compute-sanitizer is closed on the GPU pool; the thread kernels' code is plain scalar C++, so it is run here
    under AddressSanitizer + UndefinedBehaviorSanitizer instead (workspace indices, shift counts, class-list spill)."""
    import sys
    so = os.path.join(str(tmp_path), "libetc_asan.so")
    src = os.path.join(ROOT, "tests", "native", "exact_thread_check.cpp")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined",
                           "-fno-sanitize-recover=undefined", "-Wno-unknown-pragmas", "-o", so, src])
    asan = subprocess.check_output(["g++", "-print-file-name=libasan.so"]).decode().strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "exact_thread_sanitize.py"), so], env=env,
                       capture_output=True, timeout=900)
    assert r.returncode == 0 and b"sanitizer run clean" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])

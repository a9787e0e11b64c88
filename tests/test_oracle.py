"""Pins the plain-C oracle (oracle/trew_oracle.c) before anything is allowed to trust it:
(a) the known-answer vectors of the reference's own test/test.cpp, (b) golden fixtures produced by
the compiled reference (tools/make_golden.py), (c) live differential fuzzing against the compiled
reference where oracle/_ref exists.  CPU only."""
import random

import pytest

from oracle.oracle import Oracle, Reference, reference_available, seq_to_str, str_to_seq
from trew_b200 import synth


def tables_from_json(rows):
    return {(tb, k, str_to_seq(s)): c for tb, k, s, c in rows}


# ---- (a) reference test/test.cpp known-answer vectors ------------------------------------------

@pytest.mark.parametrize("bef,aft", [("ATATATTTT", "TTTTATATA"), ("GCGACTTGACGC", "TTGACGCGCGAC"),
                                     ("GGGGGGGTGGG", "TGGGGGGGGGG")])
def test_get_rot_seq_vectors(bef, aft):
    # test/test.cpp:83-97
    assert Oracle().canon(str_to_seq(bef), len(bef)) == str_to_seq(aft)


@pytest.mark.parametrize("s", ["ATTTTTTT", "ATTTTTTTGC", "ATTATAGCGATCGTCACCATTGC"])
def test_get_repeat_check_vectors(s):
    # test/test.cpp:99-109
    assert Oracle().homo(str_to_seq(s), len(s)) is False


def test_homopolymer_is_vetoed():
    o = Oracle()
    for k in (3, 17, 32, 33, 64):
        for c in range(4):
            assert o.homo(int("".join(format(c, "02b") for _ in range(k)), 2), k)


@pytest.mark.parametrize("unit", ["TTGCATCACACCCTCGCCG", "TTAGGG", "TTAGAGCCCACA",
                                  "TTTTGCCCTCATCACACCCTCGCCTCCTTCGC"])
def test_k_mer_check_perfect_repeat_64(unit):
    # test/test.cpp:172-214: unit x 20 -> exactly one entry in the high map, k == len, count == len*19+1,
    # and the RC-fold of the key is the RC-fold of the unit
    o = Oracle(5, 32)
    buf = (unit * 20).encode()
    th, tl, sh, sl, em = o.k_mer_check(buf, 0, len(buf) - 1, 5, 32)
    high = {k: v for k, v in em.items() if k[0] == 0}
    assert len(high) == 1
    (_, k, seq), cnt = next(iter(high.items()))
    assert k == len(unit) == th
    assert cnt == len(unit) * 19 + 1
    x = str_to_seq(unit)
    assert min(seq, o.crc(seq, k)) == min(x, o.crc(x, k))


@pytest.mark.parametrize("unit", ["TGCAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", "TTAGGG", "TTAGAGCCCACA",
                                  "TTTTGCCCTCATCACACCCTCGCCTCCTTCGC",
                                  "TTTTGCCCTCATCACACCCTCGCCTCCTTCGTGCTTGCCCCCACACTGACTGACGTGCAGTCTG"])
def test_k_mer_check_perfect_repeat_128(unit):
    # test/test.cpp:216-258
    o = Oracle(5, 64)
    buf = (unit * 10).encode()
    th, tl, sh, sl, em = o.k_mer_check(buf, 0, len(buf) - 1, 5, 64)
    high = {k: v for k, v in em.items() if k[0] == 0}
    assert len(high) == 1
    (_, k, seq), cnt = next(iter(high.items()))
    assert k == len(unit)
    assert cnt == len(unit) * 9 + 1
    x = str_to_seq(unit)
    assert min(seq, o.crc(seq, k)) == min(x, o.crc(x, k))


# ---- (b) golden fixtures from the compiled reference --------------------------------------------

def test_scan_golden(scan_cases):
    assert len(scan_cases) >= 15
    for case in scan_cases:
        o = Oracle(case["min_mer"], case["max_mer"], case["low"], case["high"], case["slice_len"])
        r1 = [s.encode() for s in case["reads1"]]
        r2 = [s.encode() for s in case["reads2"]] if case["reads2"] is not None else None
        got = o.scan(case["mode"], r1, r2)
        assert got == tables_from_json(case["tables"]), case["name"]


def test_kmer_check_golden(kmer_check_cases):
    for c in kmer_check_cases:
        o = Oracle(c["min_mer"], c["max_mer"])
        th, tl, sh, sl, em = o.k_mer_check(c["seq"].encode(), c["st"], c["nd"], c["kmin"], c["kmax"])
        assert (th, tl) == (c["th"], c["tl"])
        assert (seq_to_str(sh, th) if th else "") == c["Sh"]
        assert (seq_to_str(sl, tl) if tl else "") == c["Sl"]
        assert em == tables_from_json(c["emissions"])


# ---- (c) live differential fuzz against the compiled reference ---------------------------------

needs_ref = pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built (no reference tree)")


@needs_ref
@pytest.mark.parametrize("mn,mx", [(5, 32), (3, 64), (6, 33), (9, 18)])
def test_live_short(mn, mx):
    reads = synth.adversarial_short(1000 + mn * 100 + mx, 250, max_unit=mx)
    assert Oracle(mn, mx).scan(0, reads) == Reference(mn, mx, table_max_mer=max(8, mn)).scan(0, reads)


@needs_ref
@pytest.mark.parametrize("mn,mx,rl", [(5, 32, 150), (5, 32, 90), (4, 48, 100)])
def test_live_pair(mn, mx, rl):
    r1, r2 = synth.adversarial_pairs(2000 + mx + rl, 150, read_len=rl, max_unit=mx, truncate_mate2=0.1)
    # the 64-bit paired path leaks its temp map; the oracle can emulate it
    a = Oracle(mn, mx, emulate_pair_leak=True).scan(1, r1, r2)
    assert a == Reference(mn, mx, table_max_mer=8).scan(1, r1, r2)


@needs_ref
@pytest.mark.parametrize("mn,mx,sl", [(5, 32, 150), (3, 64, 128), (5, 20, 64)])
def test_live_long(mn, mx, sl):
    reads = synth.adversarial_long(3000 + sl, 40, min_len=sl - 10, max_len=1800, max_unit=mx)
    reads = [r for r in reads if len(r) >= sl]  # the reader drops shorter reads (src/kmer.cpp:1184)
    assert Oracle(mn, mx, slice_len=sl).scan(2, reads) == \
        Reference(mn, mx, slice_len=sl, table_max_mer=8).scan(2, reads)


@needs_ref
def test_live_primitives():
    rng = random.Random(5)
    o, r = Oracle(3, 64), Reference(3, 64, table_max_mer=8)
    for _ in range(2000):
        k = rng.randint(3, 64)
        v = rng.getrandbits(2 * k)
        if rng.random() < 0.3:  # periodic words: ties between rotations
            u = rng.randint(1, k)
            unit = rng.getrandbits(2 * u)
            v = 0
            for i in range(k):
                v = (v << 2) | ((unit >> (2 * (i % u))) & 3)
        assert o.canon(v, k) == r.canon(v, k)
        assert o.crc(v, k) == r.crc(v, k)
        assert o.homo(v, k) == r.homo(v, k)

"""Synthetic read generators for parity tests and benchmarks (no real data is available offline).

Two families:

* ``adversarial_*``: small mixed sets built to hit every routing branch of the reference's
  buffer_task / buffer_task_pair / buffer_task_long: perfect and noisy
  repeats of every unit length, half-repeat reads, reverse complements, indels, N runs, lowercase,
  and read lengths that flip each guard (n < 2*MIN, n < 4*MIN, n < 4*MAX).
* ``config_*``: the BASELINE.json workload shapes (fixed-length 150 bp short reads with ~1 %
  TTAGGG-repeat reads, paired fragments, HiFi-like long reads), vectorised with numpy so that tens
  of millions of reads can be generated on the GPU box in seconds.
"""
from __future__ import annotations

import random
from typing import List, Tuple

import numpy as np

BASES = b"ACGT"
COMP = bytes.maketrans(b"ACGTacgtN", b"TGCAtgcaN")
KNOWN_UNITS = [b"TTAGGG", b"TTTAGGG", b"TTAGG", b"TG", b"A", b"TTGGGG", b"TTAGGC", b"CCCTAA", b"AATGG"]


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def _rand_seq(rng: random.Random, n: int) -> bytes:
    return bytes(rng.choice(BASES) for _ in range(n))


def _repeat(rng: random.Random, unit: bytes, n: int) -> bytes:
    phase = rng.randrange(len(unit))
    reps = (n + phase) // len(unit) + 2
    return (unit * reps)[phase:phase + n]


def _mutate(rng: random.Random, s: bytes, sub: float, indel: bool) -> bytes:
    b = bytearray(s)
    if sub > 0:
        for i in range(len(b)):
            if rng.random() < sub:
                b[i] = rng.choice(BASES)
    if indel and len(b) > 10:
        p = rng.randrange(2, len(b) - 2)
        if rng.random() < 0.5:
            del b[p]
            b.append(rng.choice(BASES))
        else:
            b.insert(p, rng.choice(BASES))
            b.pop()
    return bytes(b)


def _decorate(rng: random.Random, s: bytes) -> bytes:
    b = bytearray(s)
    if rng.random() < 0.3:  # 2 % N in 30 % of reads
        for i in range(len(b)):
            if rng.random() < 0.02:
                b[i] = ord("N")
    if rng.random() < 0.1:
        b = bytearray(bytes(b).lower())
    return bytes(b)


def _unit(rng: random.Random, max_unit: int) -> bytes:
    if rng.random() < 0.4:
        return rng.choice(KNOWN_UNITS)
    return _rand_seq(rng, rng.randint(1, max_unit))


def adversarial_read(rng: random.Random, n: int, max_unit: int = 32) -> bytes:
    """One read of length n drawn from the Appendix-C mixture."""
    sub = rng.choice([0, 0, 0.01, 0.03, 0.08, 0.15, 0.30])
    kind = rng.random()
    if kind < 0.15:
        s = _rand_seq(rng, n)
    elif kind < 0.50:
        s = _mutate(rng, _repeat(rng, _unit(rng, max_unit), n), sub, rng.random() < 0.2)
    elif kind < 0.62:  # repeat in the left half only
        h = n // 2
        s = _mutate(rng, _repeat(rng, _unit(rng, max_unit), h), sub, False) + _rand_seq(rng, n - h)
    elif kind < 0.74:  # right half only
        h = n // 2
        s = _rand_seq(rng, h) + _mutate(rng, _repeat(rng, _unit(rng, max_unit), n - h), sub, False)
    elif kind < 0.86:  # different units in the two halves
        h = n // 2
        s = _mutate(rng, _repeat(rng, _unit(rng, max_unit), h), sub, False) + \
            _mutate(rng, _repeat(rng, _unit(rng, max_unit), n - h), sub, False)
    else:  # repeat occupying a random stretch
        a = rng.randrange(0, max(1, n // 2))
        b = rng.randrange(a, n)
        s = _rand_seq(rng, a) + _repeat(rng, _unit(rng, max_unit), b - a) + _rand_seq(rng, n - b)
    if rng.random() < 0.5:
        s = revcomp(s)
    return _decorate(rng, s)


SHORT_LENGTHS = [9, 12, 25, 40, 60, 75, 100, 150, 151, 246, 300]


def adversarial_short(seed: int, count: int, max_unit: int = 32, lengths=None) -> List[bytes]:
    rng = random.Random(seed)
    lengths = lengths or SHORT_LENGTHS
    return [adversarial_read(rng, rng.choice(lengths), max_unit) for _ in range(count)]


def adversarial_pairs(seed: int, count: int, read_len: int = 150, max_unit: int = 32,
                      truncate_mate2: float = 0.0) -> Tuple[List[bytes], List[bytes]]:
    """Mates cut from one fragment of length L..3L whose repeat covers a random prefix / suffix /
    middle / everything.  Mate 2 is the reverse complement of the fragment's tail."""
    rng = random.Random(seed)
    r1: List[bytes] = []
    r2: List[bytes] = []
    for _ in range(count):
        flen = rng.randint(read_len, 3 * read_len)
        kind = rng.random()
        unit = _unit(rng, max_unit)
        sub = rng.choice([0, 0, 0.01, 0.03, 0.08])
        if kind < 0.2:
            frag = _rand_seq(rng, flen)
        elif kind < 0.5:
            frag = _mutate(rng, _repeat(rng, unit, flen), sub, False)
        else:
            a = rng.randrange(0, flen)
            b = rng.randrange(a, flen + 1)
            if rng.random() < 0.5:
                a = 0
            if rng.random() < 0.5:
                b = flen
            frag = _rand_seq(rng, a) + _mutate(rng, _repeat(rng, unit, b - a), sub, False) + _rand_seq(rng, flen - b)
        if rng.random() < 0.5:
            frag = revcomp(frag)
        m1 = frag[:read_len]
        m2 = revcomp(frag[-read_len:])
        if rng.random() < truncate_mate2:
            m2 = m2[:rng.randint(max(1, read_len // 3), read_len)]
        r1.append(_decorate(rng, m1))
        r2.append(_decorate(rng, m2))
    return r1, r2


def adversarial_long(seed: int, count: int, min_len: int = 140, max_len: int = 2500,
                     max_unit: int = 32) -> List[bytes]:
    """Long reads with a repeat at the 5' end, 3' end, both ends, a unit switch or a strand switch."""
    rng = random.Random(seed)
    out: List[bytes] = []
    for _ in range(count):
        n = rng.randint(min_len, max_len)
        unit = _unit(rng, max_unit)
        sub = rng.choice([0, 0, 0.001, 0.01, 0.03])
        kind = rng.random()
        a = rng.randint(0, n)
        if kind < 0.15:
            s = _rand_seq(rng, n)
        elif kind < 0.35:
            s = _mutate(rng, _repeat(rng, unit, a), sub, False) + _rand_seq(rng, n - a)
        elif kind < 0.55:
            s = _rand_seq(rng, n - a) + _mutate(rng, _repeat(rng, unit, a), sub, False)
        elif kind < 0.70:
            a = rng.randint(0, n // 2)
            b = rng.randint(0, n // 2)
            s = _mutate(rng, _repeat(rng, unit, a), sub, False) + _rand_seq(rng, n - a - b) + \
                _mutate(rng, _repeat(rng, unit, b), sub, False)
        elif kind < 0.80:
            s = _mutate(rng, _repeat(rng, unit, n), sub, False)
        elif kind < 0.90:  # unit switch
            s = _repeat(rng, unit, a) + _repeat(rng, _unit(rng, max_unit), n - a)
        else:  # strand switch
            s = _repeat(rng, unit, a) + revcomp(_repeat(rng, unit, n - a))
        if rng.random() < 0.5:
            s = revcomp(s)
        out.append(_decorate(rng, s))
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE.json workload shapes, vectorised.
# ---------------------------------------------------------------------------------------------

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP_IDX = np.array([3, 2, 1, 0], dtype=np.uint8)  # A<->T, C<->G on indices into "ACGT"


def _telomeric_rows(rng: np.random.Generator, n: int, length: int, sub: float) -> np.ndarray:
    """n rows of (TTAGGG)^m at uniform random phase, 50 % reverse-complemented, per-base substitutions."""
    unit = np.array([3, 3, 0, 2, 2, 2], dtype=np.uint8)  # TTAGGG as indices into "ACGT"
    phase = rng.integers(0, 6, size=(n, 1))
    idx = (np.arange(length)[None, :] + phase) % 6
    rows = unit[idx]
    rc = rng.random(n) < 0.5
    rows[rc] = _COMP_IDX[rows[rc]][:, ::-1]
    if sub > 0:
        m = rng.random((n, length)) < sub
        rows[m] = rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)
    return rows


def config_short(seed: int, n_reads: int, length: int = 150, telomeric: float = 0.01,
                 half_telomeric: float = 0.002, n_rate: float = 0.001, sub: float = 0.01) -> np.ndarray:
    """BASELINE.json configs[1] shape (SURVEY.md 8(d) cfg 2): n_reads x length ASCII matrix.
    99 % i.i.d. uniform ACGT; ``telomeric`` (TTAGGG)^n reads; ``half_telomeric`` reads telomeric in one
    half only; ``n_rate`` of all bases replaced by N."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(n_reads, length), dtype=np.uint8)
    kind = rng.random(n_reads)
    tel = np.nonzero(kind < telomeric)[0]
    if tel.size:
        idx[tel] = _telomeric_rows(rng, tel.size, length, sub)
    half = np.nonzero((kind >= telomeric) & (kind < telomeric + half_telomeric))[0]
    if half.size:
        rows = _telomeric_rows(rng, half.size, length, sub)
        left = rng.random(half.size) < 0.5
        h = length // 2
        keep = idx[half]
        keep[left, :h] = rows[left, :h]
        keep[~left, h:] = rows[~left, h:]
        idx[half] = keep
    out = _ASCII[idx]
    if n_rate > 0:
        cnt = rng.binomial(n_reads * length, n_rate)
        pos = rng.integers(0, n_reads * length, size=cnt)
        out.reshape(-1)[pos] = ord("N")
    return out


def config_pairs(seed: int, n_pairs: int, length: int = 150, telomeric: float = 0.01,
                 sub: float = 0.01) -> Tuple[np.ndarray, np.ndarray]:
    """BASELINE.json configs[2] shape: mate 1 and mate 2 from one fragment; telomeric fragments give two
    repeat mates (mate 2 reverse-complemented), the rest are independent random mates."""
    rng = np.random.default_rng(seed)
    i1 = rng.integers(0, 4, size=(n_pairs, length), dtype=np.uint8)
    i2 = rng.integers(0, 4, size=(n_pairs, length), dtype=np.uint8)
    tel = np.nonzero(rng.random(n_pairs) < telomeric)[0]
    if tel.size:
        frag = _telomeric_rows(rng, tel.size, 2 * length + 50, 0.0)
        m1 = frag[:, :length].copy()
        m2 = _COMP_IDX[frag[:, -length:]][:, ::-1].copy()
        for m in (m1, m2):
            s = rng.random(m.shape) < sub
            m[s] = rng.integers(0, 4, size=int(s.sum()), dtype=np.uint8)
        i1[tel] = m1
        i2[tel] = m2
    return _ASCII[i1], _ASCII[i2]


def config_long(seed: int, n_reads: int, mean_len: int = 15000, sd_len: int = 2000, min_len: int = 1000,
                telomeric: float = 0.02, err: float = 0.001) -> List[np.ndarray]:
    """BASELINE.json configs[3] shape: HiFi-like reads, ``telomeric`` of them carry (TTAGGG)^n on the
    first or last 0.5-5 kb."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(min_len, rng.normal(mean_len, sd_len, size=n_reads).astype(np.int64))
    out: List[np.ndarray] = []
    for n in lens:
        n = int(n)
        idx = rng.integers(0, 4, size=n, dtype=np.uint8)
        if rng.random() < telomeric:
            tl = int(min(n, rng.integers(500, 5001)))
            row = _telomeric_rows(rng, 1, tl, err)[0]
            if rng.random() < 0.5:
                idx[:tl] = row
            else:
                idx[n - tl:] = row
        out.append(_ASCII[idx])
    return out


def fastq_bytes(reads, name_prefix: str = "r") -> bytes:
    """Minimal 4-line FASTQ for a list of byte strings / uint8 rows."""
    parts = []
    for i, r in enumerate(reads):
        s = bytes(r) if not isinstance(r, (bytes, bytearray)) else r
        parts.append(b"@%s%d\n%s\n+\n%s\n" % (name_prefix.encode(), i, s, b"I" * len(s)))
    return b"".join(parts)


def fastq_matrix_bytes(mat: np.ndarray) -> bytes:
    """Vectorised FASTQ for a fixed-length ASCII matrix (header '@r', quality 'I' * L)."""
    n, L = mat.shape
    rec = np.empty((n, 3 + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
    rec[:, 3:3 + L] = mat
    rec[:, 3 + L] = 10
    rec[:, 4 + L:6 + L] = np.frombuffer(b"+\n", dtype=np.uint8)
    rec[:, 6 + L:6 + 2 * L] = ord("I")
    rec[:, 6 + 2 * L] = 10
    return rec.tobytes()


# ---------------------------------------------------------------------------------------------
# numpy mirror of the device-side generator (trew_synth_resident / synth_kernel in scan_kernels.cu)
# ---------------------------------------------------------------------------------------------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _synth_hash(seed: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(0x9e3779b97f4a7c15) * (a.astype(np.uint64) + np.uint64(1))
             + np.uint64(0xbf58476d1ce4e5b9) * (b.astype(np.uint64) + np.uint64(1)))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(32)).astype(np.uint32)


def device_mirror(seed: int, n_reads: int, read_len: int = 150, tel_ppm: int = 10000, half_ppm: int = 2000,
                  n_ppm: int = 1000, sub_ppm: int = 10000, flavor: int = 0) -> np.ndarray:
    """The exact reads trew_synth_resident(_ex)() generates on the GPU, as an n_reads x read_len ASCII matrix
    (flavor 1: rows 2u, 2u+1 are the mates of pair u; flavor 2: long reads with telomeric ends)."""
    thr = lambda ppm: np.uint64((ppm << 32) // 1000000)
    L = read_len
    r = np.repeat(np.arange(n_reads, dtype=np.uint64), L).reshape(n_reads, L)
    j = np.tile(np.arange(L, dtype=np.uint64), n_reads).reshape(n_reads, L)
    frag = (r[:, :1] >> np.uint64(1)) if flavor == 1 else r[:, :1]
    kind = _synth_hash(seed, frag, np.full((n_reads, 1), 0xffffffff, dtype=np.uint64)).astype(np.uint64)
    aux = _synth_hash(seed, frag, np.full((n_reads, 1), 0xfffffffe, dtype=np.uint64))
    code = _synth_hash(seed, r, j) & np.uint32(3)
    tel = np.broadcast_to(kind < thr(tel_ppm), (n_reads, L)).copy()
    halfk = (~(kind < thr(tel_ppm))) & (kind < thr(tel_ppm) + thr(half_ppm))
    left = ((aux >> np.uint32(8)) & np.uint32(1)) == 1
    in_half = np.where(left, j < np.uint64(L // 2), j >= np.uint64(L // 2))
    tel = np.where(halfk, in_half, tel)
    if flavor == 2:
        tl = np.minimum(np.uint64(500) + (_synth_hash(seed, frag, np.full((n_reads, 1), 0xfffffffd, dtype=np.uint64)) % np.uint32(4501)).astype(np.uint64),
                        np.uint64(L))
        front = ((aux >> np.uint32(9)) & np.uint32(1)) == 1
        ends = np.where(front, j < tl, j >= np.uint64(L) - tl)
        tel = np.where(tel, ends, tel)
    phase = (aux % np.uint32(6)).astype(np.uint64)
    rc = ((aux >> np.uint32(4)) & np.uint32(1)) == 1
    if flavor == 1:
        rc = rc ^ ((r[:, :1] & np.uint64(1)) == 1)
    idx = ((j + phase) % np.uint64(6)).astype(np.int64)
    unit_f = np.array([0, 0, 3, 1, 1, 1], dtype=np.uint32)
    unit_r = np.array([2, 2, 2, 0, 3, 3], dtype=np.uint32)
    c = np.where(rc, unit_r[idx], unit_f[idx])
    sh = _synth_hash(seed ^ 0x5555555555555555, r, j)
    tcode = np.where(sh.astype(np.uint64) < thr(sub_ppm), (sh >> np.uint32(3)) & np.uint32(3), c)
    code = np.where(tel, tcode, code)
    inval = _synth_hash(seed ^ 0xaaaaaaaaaaaaaaaa, r, j).astype(np.uint64) < thr(n_ppm)
    letters = np.frombuffer(b"TGCA", dtype=np.uint8)[code.astype(np.int64)]
    return np.where(inval, np.uint8(ord("N")), letters).astype(np.uint8)


def bgzf_write(src_path: str, dst_path: str, level: int = 1, block: int = 65280) -> None:
    """Re-write a file as BGZF (bgzip): independent gzip members of at most 64 KiB with the 'BC' extra field,
    terminated by the empty EOF member.  Benchmark / test tooling."""
    import struct
    import zlib
    with open(src_path, "rb") as src, open(dst_path, "wb") as dst:
        while True:
            chunk = src.read(block)
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = c.compress(chunk) + c.flush()
            dst.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 12 + 6 + len(body) + 8 - 1))
            dst.write(body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
            if not chunk:
                break

"""Randomised differential test of the FASTQ reader: every input kind (plain mapped / plain read, gzip through the
own decoder / through zlib with random member cuts and trailing bytes, BGZF), random block sizes and parallel
thresholds, single and paired, against a Python restatement of the reference's record rules.  CPU only."""
import gzip
import os
import random

import pytest

from trew_b200 import api, synth

ENVS = ({}, {"TREW_NO_MMAP": "1", "TREW_ZLIB_GZ": "1"})


def set_env(monkeypatch, env):
    for k in ("TREW_NO_MMAP", "TREW_ZLIB_GZ", "TREW_INGEST_PAR_MIN"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)


@pytest.mark.parametrize("seed", [1, 2])
def test_single_file_soak(tmp_path, monkeypatch, seed):
    rnd = random.Random(seed)
    d = str(tmp_path)

    def line(maxlen):
        n = rnd.choice([0, 1, 2, 63, 64, 65, rnd.randrange(0, maxlen)])
        return bytes(rnd.choice(b"ACGTNacgtn\r.") for _ in range(n))

    for it in range(60):
        mode = rnd.choice([api.MODE_SHORT, api.MODE_LONG])
        sl = rnd.choice([50, 150])
        lines = []
        for _ in range(rnd.randrange(0, 300)):
            lines += [b"@" + line(40), line(300), b"+" + line(3), line(300)]
        data = b"\n".join(lines) + (b"\n" if rnd.random() < 0.8 and lines else b"")
        if rnd.random() < 0.3:
            data += b"@trailing\nACGT"
        p = os.path.join(d, "s.fastq")
        open(p, "wb").write(data)
        cuts = sorted(rnd.sample(range(len(data) + 1), min(len(data) + 1, rnd.randrange(0, 4))))
        parts = [data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])]
        gz = p + ".gz"
        open(gz, "wb").write(b"".join(gzip.compress(x, rnd.choice([0, 1, 6, 9])) for x in parts) + (b"garbage" if rnd.random() < 0.2 else b""))
        bgz = os.path.join(d, "s.fastq.bgz")
        synth.bgzf_write(p, bgz)
        seqs = data.split(b"\n")[:-1][1::4]
        want = [s for s in seqs if len(s) >= sl] if mode == api.MODE_LONG else seqs
        too_long = mode == api.MODE_SHORT and any(len(s) > 1000 for s in seqs)
        for path in (p, gz, bgz):
            for env in ENVS + ({"TREW_INGEST_PAR_MIN": str(rnd.choice([1, 100, 5000]))},):
                set_env(monkeypatch, env)
                chunk = rnd.choice([0, 64, 1000, 4096, 70000])
                rc, msg, got, _ = api.ingest_records(mode, path, slice_length=sl, chunk_bytes=chunk)
                if too_long:
                    assert rc != 0
                else:
                    assert rc == 0 and got == want, (it, os.path.basename(path), env, chunk, rc, msg, len(got), len(want))


def test_paired_files_soak(tmp_path, monkeypatch):
    rnd = random.Random(7)
    d = str(tmp_path)

    def fq(n, maxlen, hdr):
        out = []
        for _ in range(n):
            L = rnd.randrange(0, maxlen)
            s = bytes(rnd.choice(b"ACGTN") for _ in range(L))
            out.append((b"@" + b"h" * rnd.randrange(0, hdr) + b"\n" + s + b"\n+\n" + b"I" * L + b"\n", s))
        return out

    for it in range(40):
        n = rnd.randrange(0, 300)
        a = fq(n, rnd.choice([20, 300]), rnd.choice([1, 200]))      # different record sizes: the blocks of the two
        b = fq(n + (rnd.choice([1, 5]) if rnd.random() < 0.2 else 0), rnd.choice([20, 300]), rnd.choice([1, 200]))   # files drift apart
        da, db = b"".join(x for x, _ in a), b"".join(x for x, _ in b)
        pa, pb = os.path.join(d, "a.fastq"), os.path.join(d, "b.fastq")
        open(pa, "wb").write(da)
        open(pb, "wb").write(db)
        ga, gb = pa + ".gz", pb + ".gz"
        open(ga, "wb").write(gzip.compress(da, 1))
        open(gb, "wb").write(gzip.compress(db, 6))
        ba, bb = os.path.join(d, "a.bgz"), os.path.join(d, "b.bgz")
        synth.bgzf_write(pa, ba)
        synth.bgzf_write(pb, bb)
        for f1, f2 in ((pa, pb), (ga, gb), (ba, bb), (pa, gb)):
            for env in ENVS:
                set_env(monkeypatch, env)
                chunk = rnd.choice([0, 64, 500, 4096, 50000])
                rc, msg, r1, r2 = api.ingest_records(api.MODE_PAIR, f1, f2, chunk_bytes=chunk)
                if len(a) != len(b):
                    assert rc != 0 and "Mismatched" in msg, (rc, msg)
                else:
                    assert rc == 0 and r1 == [s for _, s in a] and r2 == [s for _, s in b], (it, f1, f2, env, chunk, rc, msg)

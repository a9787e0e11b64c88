#!/usr/bin/env python3
"""Generate tests/golden/*.json.gz by running the COMPILED REFERENCE (oracle/_ref) on seeded synthetic
inputs.  Run in the build container (needs /root/reference to have been compiled by
`make -C oracle`); the fixtures it writes are committed and travel to the GPU box, where the
reference tree does not exist.

    python tools/make_golden.py

Fixtures:
  scan_cases.json.gz   chunk-level cases: reads in, six count tables out (buffer_task*,
                    src/kmer.cpp:80-985), for short / paired / long mode and several
                    (MIN_MER, MAX_MER, LOW, HIGH) settings.
  kmer_check.json.gz   primitive-level cases: k_mer_check(_128) return values + emissions
                    (src/kmer.cpp:2144-2547).
  cli_cases.json.gz    whole-program cases: FASTQ text in, the reference's stdout out (H/L report of
                    process_output, src/kmer.cpp:1478-1634, and >Putative_TRM, :2571-2761).
"""
import gzip
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import REF_BIN, Reference, seq_to_str  # noqa: E402
from trew_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dump(obj, name):
    with gzip.GzipFile(os.path.join(OUT, name), "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode())


def tables_json(t):
    return [[tb, k, seq_to_str(s, k), c] for (tb, k, s), c in sorted(t.items())]


def scan_cases():
    cases = []

    def add(name, mode, mn, mx, low, high, sl, r1, r2=None):
        ref = Reference(mn, mx, low, high, sl, table_max_mer=max(8, mn))
        t = ref.scan(mode, r1, r2)
        cases.append({"name": name, "mode": mode, "min_mer": mn, "max_mer": mx, "low": low, "high": high,
                      "slice_len": sl, "reads1": [r.decode() for r in r1],
                      "reads2": [r.decode() for r in r2] if r2 is not None else None,
                      "tables": tables_json(t)})
        print("%-28s %4d reads -> %5d entries" % (name, len(r1), len(t)))

    add("short_5_32", 0, 5, 32, 0.5, 0.8, 150, synth.adversarial_short(1, 220))
    add("short_3_64", 0, 3, 64, 0.5, 0.8, 150, synth.adversarial_short(2, 160, max_unit=64))
    add("short_5_64", 0, 5, 64, 0.5, 0.8, 150, synth.adversarial_short(3, 120, max_unit=64))
    add("short_7_20", 0, 7, 20, 0.5, 0.8, 150, synth.adversarial_short(4, 160, max_unit=20))
    add("short_12_40", 0, 12, 40, 0.5, 0.8, 150, synth.adversarial_short(5, 120, max_unit=40))
    add("short_5_32_L03_H06", 0, 5, 32, 0.3, 0.6, 150, synth.adversarial_short(6, 120))
    add("short_5_32_L08_H08", 0, 5, 32, 0.8, 0.8, 150, synth.adversarial_short(7, 120))
    add("short_5_32_L1_H1", 0, 5, 32, 1.0, 1.0, 150, synth.adversarial_short(8, 120))
    # paired: at 150 bp / MAX 32 the large-k block (and the 64-bit clear() bug) is unreachable
    r1, r2 = synth.adversarial_pairs(9, 160, read_len=150)
    add("pair_5_32_150bp", 1, 5, 32, 0.5, 0.8, 150, r1, r2)
    # 128-bit paired path (clears its temp map), short mates with truncation -> large-k block
    r1, r2 = synth.adversarial_pairs(10, 160, read_len=100, max_unit=40, truncate_mate2=0.15)
    add("pair_5_40_100bp_trunc", 1, 5, 40, 0.5, 0.8, 150, r1, r2)
    r1, r2 = synth.adversarial_pairs(11, 100, read_len=120, max_unit=64, truncate_mate2=0.1)
    add("pair_3_64_120bp_trunc", 1, 3, 64, 0.5, 0.8, 150, r1, r2)
    add("long_5_32", 2, 5, 32, 0.5, 0.8, 150, synth.adversarial_long(12, 50))
    add("long_3_64_s140", 2, 3, 64, 0.5, 0.8, 140, synth.adversarial_long(13, 40, max_unit=64))
    add("long_5_32_s100", 2, 5, 32, 0.5, 0.8, 100, synth.adversarial_long(14, 40, min_len=90, max_len=1500))
    # the north-star shape, small: 150 bp, ~TTAGGG reads, N's
    mat = synth.config_short(15, 600, telomeric=0.10, half_telomeric=0.05, n_rate=0.002)
    add("cfg2_shape_small", 0, 5, 32, 0.5, 0.8, 150, [bytes(r) for r in mat])
    dump(cases, "scan_cases.json.gz")


def kmer_check_cases():
    rng = random.Random(99)
    cases = []
    for (mn, mx) in [(5, 32), (5, 64), (3, 20)]:
        ref = Reference(mn, mx, table_max_mer=8)
        for i in range(40):
            n = rng.choice([30, 75, 150, 299])
            s = synth.adversarial_read(rng, n, max_unit=mx)
            st = rng.randrange(0, 5)
            nd = n - 1 - rng.randrange(0, 5)
            kmax = min(mx, max(mn, (nd - st + 1) // 2))
            th, tl, sh, sl_, em = ref.k_mer_check(s, st, nd, mn, kmax)
            cases.append({"min_mer": mn, "max_mer": mx, "seq": s.decode(), "st": st, "nd": nd, "kmin": mn,
                          "kmax": kmax, "th": th, "tl": tl,
                          "Sh": seq_to_str(sh, th) if th else "", "Sl": seq_to_str(sl_, tl) if tl else "",
                          "emissions": tables_json(em)})
    dump(cases, "kmer_check.json.gz")
    print("kmer_check cases:", len(cases))


def cli_cases():
    cases = []
    with tempfile.TemporaryDirectory() as td:
        def run(name, args, files):
            paths = {}
            for fn, content in files.items():
                p = os.path.join(td, fn)
                open(p, "wb").write(content)
                paths[fn] = p
            argv = [REF_BIN] + [paths.get(a, a) for a in args]
            out = subprocess.run(argv, capture_output=True, check=True).stdout.decode()
            for fn, p in paths.items():
                out = out.replace(os.path.realpath(p), "<" + fn + ">")
            # FASTQ text is rebuilt in the tests from the sequences alone (synth.fastq_bytes layout)
            cases.append({"name": name, "args": args,
                          "files": {k: v.decode().split("\n")[1::4] for k, v in files.items()},
                          "stdout": out})
            print("%-24s %d lines" % (name, out.count("\n")))

        mat = synth.config_short(21, 1200, telomeric=0.05, half_telomeric=0.03, n_rate=0.001)
        run("short_5_32", ["short", "5", "32", "a.fastq", "-t", "2"], {"a.fastq": synth.fastq_bytes(mat)})
        mat2 = synth.config_short(22, 800, telomeric=0.04, half_telomeric=0.02)
        run("short_two_files", ["short", "5", "32", "a.fastq", "b.fastq", "-t", "3"],
            {"a.fastq": synth.fastq_bytes(mat), "b.fastq": synth.fastq_bytes(mat2)})
        run("short_7_20_adv", ["short", "7", "20", "a.fastq", "-t", "2"],
            {"a.fastq": synth.fastq_bytes(synth.adversarial_short(23, 700, max_unit=20, lengths=[100, 150, 151, 246]))})
        run("short_3_64_adv", ["short", "3", "64", "a.fastq", "-t", "2"],
            {"a.fastq": synth.fastq_bytes(synth.adversarial_short(24, 400, max_unit=64))})
        m1, m2 = synth.config_pairs(25, 800, telomeric=0.05)
        run("pair_5_32", ["short", "5", "32", "--paired_end", "--fq1", "a.fastq", "--fq2", "b.fastq", "-t", "2"],
            {"a.fastq": synth.fastq_bytes(m1), "b.fastq": synth.fastq_bytes(m2)})
        # few distinct units, well separated counts: no ties at any top-4 cut of get_score_map, so the
        # reference's >Putative_TRM is a function of its input here
        rng = random.Random(27)
        tf = []
        for unit, n_fwd, n_rev in [(b"TTAGGG", 90, 7), (b"TTTAGGG", 11, 40), (b"TTAGGC", 25, 3), (b"TGTGGG", 5, 14)]:
            for i in range(n_fwd + n_rev):
                s = synth._repeat(rng, unit, 150)
                tf.append(s if i < n_fwd else synth.revcomp(s))
        tf += [synth._rand_seq(rng, 150) for _ in range(200)]
        rng.shuffle(tf)
        run("short_tie_free", ["short", "5", "32", "a.fastq", "-t", "2"], {"a.fastq": synth.fastq_bytes(tf)})
        lr = synth.config_long(26, 40, mean_len=3000, sd_len=800, min_len=200, telomeric=0.4, err=0.002)
        run("long_5_32", ["long", "5", "32", "a.fastq", "-t", "2"], {"a.fastq": synth.fastq_bytes(lr)})
    dump(cases, "cli_cases.json.gz")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    scan_cases()
    kmer_check_cases()
    cli_cases()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

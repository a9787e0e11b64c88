"""Whole-program parity: the `trew` binary (FASTQ in, report out) against the reference's stdout captured in
tests/golden/cli_cases.json.gz, plus file-level ingest through trew_dev_process_file."""
import gzip
import os
import subprocess

import pytest

from trew_b200 import api, synth
from test_report import parse_cli_args, putative, split_sections

pytestmark = pytest.mark.gpu


def materialise(case, tmp_path, gz=False):
    paths = {}
    for name, reads in case["files"].items():
        fn = name + (".gz" if gz else "")
        p = os.path.join(tmp_path, fn)
        data = synth.fastq_bytes([r.encode() for r in reads])
        with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
            f.write(data)
        paths[name] = p
    return paths


@pytest.mark.parametrize("gz", [False, True])
def test_cli_matches_reference_stdout(cli_cases, tmp_path, gz):
    for case in cli_cases:
        paths = materialise(case, str(tmp_path), gz)
        argv = [api.CLI_PATH] + [paths.get(a, a) for a in case["args"]]
        out = subprocess.run(argv, capture_output=True, check=True).stdout.decode()
        for name, p in paths.items():
            out = out.replace(os.path.realpath(p), "<" + name + ">")
        got, want = split_sections(out), split_sections(case["stdout"])
        assert [h for h, _ in got] == [h for h, _ in want], case["name"]
        for (h, g), (_, w) in zip(got, want):
            if h != ">Putative_TRM":
                assert g == w, (case["name"], h)
        if case["name"] in ("short_tie_free", "pair_5_32"):
            assert putative(got) == putative(want)


def test_cli_argument_errors(tmp_path):
    p = os.path.join(str(tmp_path), "a.fastq")
    open(p, "wb").write(synth.fastq_bytes([b"ACGT" * 30]))
    for args, msg in [(["short", "6", "5", p], "MIN_MER must not be greater than MAX_MER."),
                      (["short", "2", "5", p], "MIN_MER must be greater than or equal to 3."),
                      (["short", "5", "65", p], "MAX_MER must be less than or equal to 64."),
                      (["long", "5", "32", p, "-s", "60"], "SLICE_LENGTH must be greater than or equal to twice of MAX_MER."),
                      (["short", "5", "32", p, "-t", "1"], "You must use at least two threads."),
                      (["short", "5", "32", p, "-L", "0.9", "-H", "0.8"], "Low baseline must be smaller than high baseline."),
                      (["short", "5", "32", "/nonexistent.fastq"], "/nonexistent.fastq : file not found")]:
        r = subprocess.run([api.CLI_PATH] + args, capture_output=True)
        assert r.returncode == 1 and msg in r.stderr.decode(), args
    long_read = os.path.join(str(tmp_path), "l.fastq")
    open(long_read, "wb").write(synth.fastq_bytes([b"ACGT" * 300]))
    r = subprocess.run([api.CLI_PATH, "short", "5", "32", long_read], capture_output=True)
    assert r.returncode == 1 and b"This mode is designed for short-read sequencing. Please use 'trew long'." in r.stderr
    assert r.stdout == b""


FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")


@pytest.mark.parametrize("args,name", [(["short", "5", "32"], "test.fastq.gz"), (["short", "5", "32"], "test.fastq"),
                                       (["short", "5", "64"], "test.fastq"), (["long", "5", "32"], "test_long.fastq.gz"),
                                       (["long", "5", "32"], "test_long.fastq"), (["long", "5", "64", "-t", "4"], "test_long.fastq.gz")])
def test_reference_fixtures_give_the_empty_skeleton(args, name):
    """BASELINE.json configs[0] through the CUDA path: the reference's own fixtures (test/test.cpp:260-443) with the
    compiled reference's stdout."""
    p = os.path.join(FIX, name)
    out = subprocess.run([api.CLI_PATH] + args + [p], capture_output=True, check=True).stdout.decode()
    rp = os.path.realpath(p)
    assert out == ">H:%s\n>L:%s\n>Putative_TRM\nNO_PUTATIVE_TRM,-1\n" % (rp, rp)


@pytest.mark.parametrize("name", ["test.fastq", "test.fastq.gz"])
def test_reference_fixture_rows_3_64(name):
    """`trew short 3 64 test/test.fastq`: the one bundled-fixture run with rows; golden = the
    compiled reference's stdout, see tests/test_fixtures.py."""
    from test_fixtures import L_ROWS_3_64, PUTATIVE_3_64
    p = os.path.join(FIX, name)
    out = subprocess.run([api.CLI_PATH, "short", "3", "64", p], capture_output=True, check=True).stdout.decode()
    rp = os.path.realpath(p)
    sec = dict(split_sections(out))
    assert sec[">H:" + rp] == []
    assert sec[">L:" + rp] == sorted(L_ROWS_3_64)
    strip = lambda rows: sorted((r.split(",")[0], r.split(",")[1], r.split(",")[3]) for r in rows)
    assert strip(sec[">Putative_TRM"]) == strip(PUTATIVE_3_64)


def test_cli_streams_sections_per_file(tmp_path):
    """Each file's >H: / >L: sections are on stdout before the next file is touched (process_output prints per file,
    src/kmer.cpp:1615-1631): a later file that fails leaves the earlier sections behind, exit code 1."""
    ok = os.path.join(FIX, "test.fastq")
    bad = os.path.join(str(tmp_path), "bad.fastq")
    open(bad, "wb").write(synth.fastq_bytes([b"ACGT" * 300]))
    r = subprocess.run([api.CLI_PATH, "short", "5", "32", ok, bad], capture_output=True)
    rp = os.path.realpath(ok)
    assert r.returncode == 1
    assert r.stdout.decode() == ">H:%s\n>L:%s\n" % (rp, rp)
    assert b"Please use 'trew long'" in r.stderr


def test_bundled_fixture_skeleton(tmp_path):
    # config 1 of BASELINE.json: non-repetitive fixtures give the empty skeleton
    p = os.path.join(str(tmp_path), "test.fastq.gz")
    reads = [bytes(r) for r in synth.config_short(5, 100, length=246, telomeric=0, half_telomeric=0, n_rate=0)]
    with gzip.open(p, "wb") as f:
        f.write(synth.fastq_bytes(reads))
    out = subprocess.run([api.CLI_PATH, "short", "5", "32", p], capture_output=True, check=True).stdout.decode()
    rp = os.path.realpath(p)
    assert out == ">H:%s\n>L:%s\n>Putative_TRM\nNO_PUTATIVE_TRM,-1\n" % (rp, rp)


def test_process_file_equals_submit(tmp_path):
    reads = synth.adversarial_short(31, 5000, lengths=[100, 150, 151])
    p = os.path.join(str(tmp_path), "a.fastq.gz")
    with gzip.open(p, "wb") as f:
        f.write(synth.fastq_bytes(reads))
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.process_file(p)
        a = ctx.finish()
        ctx.reset()
        ctx.submit_reads(reads)
        b = ctx.finish()
    assert a == b and len(a) > 0


def test_process_file_every_input_kind(tmp_path, monkeypatch):
    """Plain (mapped and read), gzip (own decoder and zlib), multi-member gzip and BGZF inputs of the same records give
    the same tables as submitting the reads directly, single and paired."""
    r1 = synth.adversarial_short(32, 4000, lengths=[100, 150, 151])
    r2 = synth.adversarial_short(33, 4000, lengths=[100, 150, 151])
    d = str(tmp_path)
    paths = {}
    for tag, reads in (("1", r1), ("2", r2)):
        data = synth.fastq_bytes(reads)
        plain = os.path.join(d, "r%s.fastq" % tag)
        open(plain, "wb").write(data)
        gz = plain + ".gz"
        with gzip.open(gz, "wb") as f:
            f.write(data)
        multi = os.path.join(d, "m%s.fastq.gz" % tag)
        cut = len(data) // 3
        cut = data.index(b"\n", cut) + 1
        open(multi, "wb").write(gzip.compress(data[:cut]) + gzip.compress(data[cut:], 1))
        bgz = os.path.join(d, "b%s.fastq.bgz" % tag)
        synth.bgzf_write(plain, bgz)
        paths[tag] = {"plain": plain, "gz": gz, "multi": multi, "bgzf": bgz}
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_reads(r1)
        want = ctx.finish()
        assert len(want) > 0
        for env in ({}, {"TREW_NO_MMAP": "1", "TREW_ZLIB_GZ": "1"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            for kind, p in paths["1"].items():
                ctx.reset()
                ctx.process_file(p)
                assert ctx.finish() == want, (kind, env)
    with api.DeviceContext(api.MODE_PAIR, 5, 32) as ctx:
        for k in ("TREW_NO_MMAP", "TREW_ZLIB_GZ"):
            monkeypatch.delenv(k, raising=False)
        ctx.submit_reads(r1, r2)
        want = ctx.finish()
        for kind in ("plain", "gz", "bgzf"):
            ctx.reset()
            ctx.process_file(paths["1"][kind], paths["2"][kind])
            assert ctx.finish() == want, kind

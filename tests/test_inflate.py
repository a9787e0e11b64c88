"""The DEFLATE decoder behind the .gz ingest path (trew_b200/csrc/inflate.cpp) against zlib.  CPU only.

tests/native/inflate_check.cpp compresses varied data with zlib's deflate (all block types, levels, strategies, flush
points), decodes it with the library's decoder whole and in arbitrary input / output chunks, and feeds it corrupted
streams; it is built with AddressSanitizer + UBSan when the toolchain has them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("inflate") / "inflate_check")
    src = [os.path.join(ROOT, "tests", "native", "inflate_check.cpp"), os.path.join(ROOT, "trew_b200", "csrc", "inflate.cpp")]
    base = ["g++", "-O1", "-g", "-std=c++17", "-o", out] + src + ["-lz", "-lpthread"]
    san = base[:4] + ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] + base[4:]
    if subprocess.run(san, capture_output=True).returncode != 0:
        r = subprocess.run(base, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    return out


@pytest.mark.parametrize("seed,no_bmi2", [(1, False), (2, False), (3, True)])
def test_decoder_matches_zlib_and_survives_corruption(checker, seed, no_bmi2):
    env = dict(os.environ, TREW_NO_BMI2="1") if no_bmi2 else dict(os.environ)   # the portable build of the decode loop
    r = subprocess.run([checker, str(seed), "14"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout[-500:], r.stderr[-2000:])


@pytest.fixture(scope="module")
def parallel_checker(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pinflate") / "pinflate_check")
    src = [os.path.join(ROOT, "tests", "native", "pinflate_check.cpp"), os.path.join(ROOT, "trew_b200", "csrc", "pinflate.cpp")]
    base = ["g++", "-O1", "-g", "-std=c++17", "-o", out] + src + ["-lz", "-lpthread"]
    san = base[:4] + ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] + base[4:]
    if subprocess.run(san, capture_output=True).returncode != 0:
        r = subprocess.run(base, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    return out


@pytest.mark.parametrize("seed,threads,segment", [(1, 4, 20000), (2, 3, 3000), (3, 8, 1 << 19)])
def test_parallel_decoder_matches_zlib_and_survives_corruption(parallel_checker, seed, threads, segment):
    """trew_b200/csrc/pinflate.cpp (speculative block starts, 16-bit symbols, chained windows) against zlib: FASTQ-like,
    incompressible, run-length and empty data at several levels / strategies / flush points, whole and with corrupted or
    truncated input, for segment sizes from a few blocks down to less than one."""
    env = dict(os.environ, TREW_PGZ_SEGMENT=str(segment))
    r = subprocess.run([parallel_checker, str(seed), str(threads)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout[-500:], r.stderr[-2000:])

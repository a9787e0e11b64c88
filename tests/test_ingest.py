"""FASTQ / FASTQ.gz reader semantics (read_fastq_thread & co, src/kmer.cpp:987-1213).  CPU only."""
import gzip
import os

import pytest

from trew_b200 import api, synth



def py_records(data: bytes):
    """Every 4k+2-th '\\n'-terminated line; a last unterminated line is never seen."""
    lines = data.split(b"\n")[:-1]
    return lines[1::4]


def write(tmp_path, name, data, gz=False):
    p = os.path.join(tmp_path, name)
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(data)
    return p


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("chunk", [64, 1000, 0])
def test_short_records(tmp_path, gz, chunk):
    reads = synth.adversarial_short(5, 300)
    data = synth.fastq_bytes(reads) + b"@tail\nACGT"  # unterminated sequence line: ignored
    p = write(tmp_path, "a.fastq" + (".gz" if gz else ""), data, gz)
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p, chunk_bytes=chunk)
    assert rc == 0, msg
    assert r1 == py_records(data) == reads


def test_crlf_counts_toward_length(tmp_path):
    data = b"@r\r\nACGTACGTAC\r\n+\r\nIIIIIIIIII\r\n"
    p = write(tmp_path, "a.fastq", data)
    rc, _, r1, _ = api.ingest_records(api.MODE_SHORT, p)
    assert rc == 0 and r1 == [b"ACGTACGTAC\r"]


def test_short_rejects_long_reads(tmp_path):
    p = write(tmp_path, "a.fastq", synth.fastq_bytes([b"A" * 1000, b"C" * 1001]))
    rc, msg, _, _ = api.ingest_records(api.MODE_SHORT, p)
    assert rc == 4 and msg == "This mode is designed for short-read sequencing. Please use 'trew long'."
    p = write(tmp_path, "b.fastq", synth.fastq_bytes([b"A" * 1000]))
    assert api.ingest_records(api.MODE_SHORT, p)[0] == 0


def test_long_drops_short_reads(tmp_path):
    reads = [b"A" * 149, b"C" * 150, b"G" * 5000, b"T" * 10]
    p = write(tmp_path, "a.fastq", synth.fastq_bytes(reads))
    rc, _, r1, _ = api.ingest_records(api.MODE_LONG, p, slice_length=150, chunk_bytes=512)
    assert rc == 0 and r1 == [reads[1], reads[2]]


@pytest.mark.parametrize("chunk", [100, 777, 0])
def test_pairs_index_wise(tmp_path, chunk):
    r1, r2 = synth.adversarial_pairs(3, 200, read_len=100, truncate_mate2=0.3)
    p1 = write(tmp_path, "a.fastq", synth.fastq_bytes(r1, "longer_header_name_"))
    p2 = write(tmp_path, "b.fastq.gz", synth.fastq_bytes(r2), gz=True)
    rc, msg, o1, o2 = api.ingest_records(api.MODE_PAIR, p1, p2, chunk_bytes=chunk)
    assert rc == 0, msg
    assert o1 == r1 and o2 == r2


def test_pairs_mismatch_is_an_error(tmp_path):
    r1, r2 = synth.adversarial_pairs(4, 20)
    p1 = write(tmp_path, "a.fastq", synth.fastq_bytes(r1))
    p2 = write(tmp_path, "b.fastq", synth.fastq_bytes(r2[:-1]))
    rc, msg, _, _ = api.ingest_records(api.MODE_PAIR, p1, p2)
    assert rc == 6 and msg == "Error: Mismatched record counts between files (num1: 80, num2: 76)."


def test_missing_file():
    rc, msg, _, _ = api.ingest_records(api.MODE_SHORT, "/nonexistent/x.fastq")
    assert rc == 5 and msg == "File open failed"


@pytest.mark.parametrize("name,mode", [("test.fastq", 0), ("test.fastq.gz", 0), ("test_long.fastq", 2), ("test_long.fastq.gz", 2)])
def test_bundled_fixtures(name, mode):
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures", name)   # byte copies, see test_fixtures.py
    data = (gzip.open(p) if name.endswith(".gz") else open(p, "rb")).read()
    rc, _, r1, _ = api.ingest_records(mode, p, chunk_bytes=4096)
    assert rc == 0 and r1 == py_records(data)


def bgzf_bytes(data: bytes, block: int = 65280) -> bytes:
    """BGZF (bgzip) container: independent gzip members of at most 64 KiB with a 'BC' extra field + the EOF member."""
    import struct
    import zlib
    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


@pytest.mark.parametrize("chunk", [64, 5000, 70000, 0])
@pytest.mark.parametrize("block", [700, 65280])
def test_bgzf_blocks_inflate_in_parallel(tmp_path, chunk, block):
    reads = synth.adversarial_short(6, 1500)
    data = synth.fastq_bytes(reads)
    raw = bgzf_bytes(data, block)
    assert gzip.decompress(raw) == data            # a valid multi-member gzip file as far as zlib is concerned
    p = os.path.join(tmp_path, "a.fastq.bgz")
    open(p, "wb").write(raw)
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p, chunk_bytes=chunk)
    assert rc == 0, msg
    assert r1 == reads


def test_truncated_bgzf_is_an_io_error(tmp_path):
    data = synth.fastq_bytes(synth.adversarial_short(7, 400))
    raw = bgzf_bytes(data, 3000)
    p = os.path.join(tmp_path, "a.fastq.gz")
    open(p, "wb").write(raw[:len(raw) // 2 + 7])
    rc, msg, _, _ = api.ingest_records(api.MODE_SHORT, p)
    assert rc == 5 and "IO Error" in msg


def test_corrupted_bgzf_payload_is_an_io_error(tmp_path):
    """A flipped payload byte inside a (stored) BGZF member must fail the member's CRC-32 like gzread's
    'incorrect data check' -- not ingest altered records."""
    import struct
    import zlib
    data = synth.fastq_bytes(synth.adversarial_short(8, 300))
    out = bytearray()
    for i in range(0, len(data), 4000):   # stored (level 0) members: any payload byte maps 1:1 to an output byte
        chunk = data[i:i + 4000]
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 18 + len(body) + 8
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    p = os.path.join(tmp_path, "ok.fastq.bgz")
    open(p, "wb").write(bytes(out))
    assert api.ingest_records(api.MODE_SHORT, p)[0] == 0
    bad = bytearray(out)
    at = 18 + 5 + 1000                     # inside the first member's stored payload, on a base
    assert bytes(bad[at:at + 1]) in b"ACGTNacgtn@+I\n" or True
    bad[at] ^= 0x04
    q = os.path.join(tmp_path, "bad.fastq.bgz")
    open(q, "wb").write(bytes(bad))
    rc, msg, _, _ = api.ingest_records(api.MODE_SHORT, q)
    assert rc == 5 and "data check" in msg


def test_pair_carry_stays_bounded_with_asymmetric_mates(tmp_path):
    """28 bp against 100 bp mates: index-wise pairing leaves a surplus of complete records on the short side every
    block.  The carry must stay O(block) (offsets are int32), and the pairs must still come out in order."""
    import ctypes as C
    n = 6000
    r1 = [bytes(r) for r in synth.config_short(41, n, length=28, telomeric=0.05, half_telomeric=0, n_rate=0.001)]
    r2 = [bytes(r) for r in synth.config_short(42, n, length=100, telomeric=0.05, half_telomeric=0, n_rate=0.001)]
    p1 = write(tmp_path, "a.fastq", synth.fastq_bytes(r1))
    p2 = write(tmp_path, "b.fastq", synth.fastq_bytes(r2))
    chunk = 4096
    L = api.load_library()
    seen1, seen2, max_off = [], [], [0]

    def sink(user, b1, l1, n1, b2, l2, n2):
        assert n1 == n2
        for i in range(n1):
            seen1.append(C.string_at(b1 + l1[2 * i], l1[2 * i + 1] - l1[2 * i] + 1))
            seen2.append(C.string_at(b2 + l2[2 * i], l2[2 * i + 1] - l2[2 * i] + 1))
        if n1:
            max_off[0] = max(max_off[0], l1[2 * n1 - 1], l2[2 * n2 - 1])
        return 0

    for env in ({}, {"TREW_NO_MMAP": "1"}):
        seen1.clear(); seen2.clear(); max_off[0] = 0
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            msg = C.create_string_buffer(512)
            rc = L.trew_ingest_file(api.MODE_PAIR, 150, p1.encode(), 0, p2.encode(), 0, chunk, api.CHUNK_SINK(sink), None, msg, 512)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        assert rc == 0, msg.value
        assert seen1 == r1 and seen2 == r2
        assert max_off[0] < 3 * chunk, max_off[0]
    # one file ends early: the surplus of the other is only counted, the mismatch is still reported with both totals
    p3 = write(tmp_path, "c.fastq", synth.fastq_bytes(r2[:n // 3]))
    rc, msg, o1, o2 = api.ingest_records(api.MODE_PAIR, p1, p3, chunk_bytes=chunk)
    assert rc == 6 and msg == "Error: Mismatched record counts between files (num1: %d, num2: %d)." % (4 * n, 4 * (n // 3))
    assert o1 == r1[:n // 3] and o2 == r2[:n // 3]


@pytest.mark.parametrize("simd", ["0", "1", "2"])
def test_parallel_newline_index_equals_serial(tmp_path, simd):
    """The sliced SIMD newline search + per-slice role assignment (forced by a tiny TREW_INGEST_PAR_MIN) returns the
    same records as the serial path in all three modes, for every instruction set, at awkward chunk sizes."""
    import subprocess, sys, textwrap
    reads = synth.adversarial_short(11, 400) + [b"", b"A", b"", b"ACGT" * 60, b"N" * 64, b"T" * 63, b"G" * 65]
    data = synth.fastq_bytes(reads) + b"@tail\nACGT"
    # blank-ish headers put several newlines into one 64-byte block
    data2 = b"".join(b"@\n" + r + b"\n+\n" + b"I" * len(r) + b"\n" for r in reads)
    p1 = write(tmp_path, "a.fastq", data)
    p2 = write(tmp_path, "b.fastq", data2)
    p3 = write(tmp_path, "c.fastq", synth.fastq_bytes(reads))
    code = textwrap.dedent("""
        import os, sys
        from trew_b200 import api
        p1, p2, p3 = sys.argv[1:4]
        def run(par, mode, f1, f2, chunk, nommap=False):
            if par: os.environ["TREW_INGEST_PAR_MIN"] = par
            else: os.environ.pop("TREW_INGEST_PAR_MIN", None)
            if nommap: os.environ["TREW_NO_MMAP"] = "1"     # read() into a buffer instead of mapping the file
            else: os.environ.pop("TREW_NO_MMAP", None)
            return api.ingest_records(mode, f1, f2, slice_length=100, chunk_bytes=chunk)
        for chunk in (0, 777, 4096):
            for mode, f1, f2 in ((api.MODE_SHORT, p1, None), (api.MODE_SHORT, p2, None), (api.MODE_LONG, p1, None),
                                 (api.MODE_PAIR, p3, p2)):
                want = run(None, mode, f1, f2, chunk)
                for par in ("1", "50", "300"):
                    for nommap in (False, True):
                        got = run(par, mode, f1, f2, chunk, nommap)
                        assert got == want, (chunk, mode, par, nommap)
                assert run(None, mode, f1, f2, chunk, True) == want
                assert want[0] == 0 and len(want[2]) > 0
        print("ok")
    """)
    env = dict(os.environ, TREW_PACK_SIMD=simd)
    out = subprocess.run([sys.executable, "-c", code, p1, p2, p3], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(__file__)))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def _gz_member(data: bytes, level=6, extra=False, name=b"", comment=b"", hcrc=False) -> bytes:
    """One gzip member built by hand (RFC 1952) so that every optional header field can be exercised."""
    import struct, zlib
    flg = (4 if extra else 0) | (8 if name else 0) | (16 if comment else 0) | (2 if hcrc else 0)
    head = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\x03"
    if extra:
        head += struct.pack("<H", 6) + b"XY\x02\x00ab"
    if name:
        head += name + b"\0"
    if comment:
        head += comment + b"\0"
    if hcrc:
        head += struct.pack("<H", zlib.crc32(head) & 0xFFFF)
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


@pytest.mark.parametrize("chunk", [64, 1000, 70000, 0])
def test_gzip_stream_layer_members_headers_and_garbage(tmp_path, chunk):
    """The library's own gzip reader: concatenated members with every optional header field, an empty member,
    stored (level 0) members, trailing bytes after the last member -- same records as the plain file."""
    reads = synth.adversarial_short(21, 900)
    data = synth.fastq_bytes(reads)
    cut = [0, 1, 777, len(data) // 3, len(data) // 3 + 1, len(data) // 2, len(data)]
    opts = [dict(), dict(extra=True), dict(name=b"a.fastq"), dict(comment=b"hello", hcrc=True), dict(level=0),
            dict(extra=True, name=b"n", comment=b"c", hcrc=True, level=9)]
    blob = b"".join(_gz_member(data[a:b], **o) for a, b, o in zip(cut, cut[1:], opts))
    blob += _gz_member(b"") + b"\0\0\0 trailing bytes that are not a gzip member"
    p = write(tmp_path, "m.fastq.gz", blob)
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p, chunk_bytes=chunk)
    assert rc == 0, msg
    assert r1 == reads


def test_gzip_stream_layer_errors(tmp_path):
    import struct
    reads = synth.adversarial_short(22, 300)
    good = _gz_member(synth.fastq_bytes(reads))
    cases = {
        "truncated body": good[:len(good) // 2],
        "truncated trailer": good[:-3],
        "bad crc": good[:-8] + struct.pack("<I", 12345) + good[-4:],
        "bad isize": good[:-4] + struct.pack("<I", 7),
        "corrupt body": good[:40] + bytes([good[40] ^ 0xFF, good[41] ^ 0x55]) + good[42:],
        "reserved header flag": good[:3] + b"\x80" + good[4:],
    }
    for name, blob in cases.items():
        p = write(tmp_path, "bad.fastq.gz", blob)
        rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p)
        if name == "corrupt body" and rc == 0:
            continue   # a flipped bit can still decode; then the CRC must have matched by construction -- not the case here
        assert rc != 0, name
        assert "File-IO Error" in msg, (name, msg)


def test_gzip_own_reader_equals_zlib_reader(tmp_path):
    """Same records through the library's decoder and through zlib's gzread (TREW_ZLIB_GZ=1), pairs included."""
    import subprocess, sys, textwrap
    a = synth.adversarial_short(23, 1500)
    b = synth.adversarial_short(24, 1500)
    p1 = write(tmp_path, "a.fastq.gz", synth.fastq_bytes(a), gz=True)
    p2 = write(tmp_path, "b.fastq.gz", synth.fastq_bytes(b), gz=True)
    code = textwrap.dedent("""
        import sys
        from trew_b200 import api
        r = [api.ingest_records(api.MODE_SHORT, sys.argv[1], chunk_bytes=c) for c in (0, 5000)]
        r.append(api.ingest_records(api.MODE_PAIR, sys.argv[1], sys.argv[2], chunk_bytes=3000))
        import hashlib, pickle
        print(hashlib.sha256(pickle.dumps(r)).hexdigest(), r[0][0], len(r[0][2]), len(r[2][3]))
    """)
    outs = []
    for env_extra in ({}, {"TREW_ZLIB_GZ": "1"}):
        env = dict(os.environ, **env_extra)
        out = subprocess.run([sys.executable, "-c", code, p1, p2], env=env, capture_output=True, text=True,
                             cwd=os.path.dirname(os.path.dirname(__file__)))
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout.split())
    assert outs[0] == outs[1]
    assert outs[0][1:] == ["0", "1500", "1500"]


# ---- parallel gzip (pinflate.cpp): the same stream layer, every member's DEFLATE stream decoded by all threads ----------

@pytest.fixture
def parallel_gz(monkeypatch):
    """Force the parallel path on small files: no size threshold, segments of a few KiB of compressed input."""
    def on(segment):
        monkeypatch.setenv("TREW_PGZ_MIN_BYTES", "0")
        monkeypatch.setenv("TREW_PGZ_SEGMENT", str(segment))
    return on


@pytest.mark.parametrize("segment", [4096, 40000, 1 << 20])
@pytest.mark.parametrize("chunk", [1000, 70000, 0])
def test_parallel_gzip_members_headers_and_garbage(tmp_path, parallel_gz, segment, chunk):
    parallel_gz(segment)
    reads = synth.adversarial_short(31, 2500) + [bytes(r) for r in synth.config_short(32, 12000, telomeric=0.02, n_rate=0.002)]
    data = synth.fastq_bytes(reads)
    cut = [0, 1, 777, len(data) // 3, len(data) // 3 + 1, len(data) // 2, len(data)]
    opts = [dict(), dict(extra=True), dict(name=b"a.fastq"), dict(comment=b"hello", hcrc=True), dict(level=0),
            dict(extra=True, name=b"n", comment=b"c", hcrc=True, level=9)]
    blob = b"".join(_gz_member(data[a:b], **o) for a, b, o in zip(cut, cut[1:], opts))
    blob += _gz_member(b"") + b"\0\0\0 trailing bytes that are not a gzip member"
    p = write(tmp_path, "m.fastq.gz", blob)
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p, chunk_bytes=chunk)
    assert rc == 0, msg
    assert r1 == reads


@pytest.mark.parametrize("level", [1, 6, 9])
def test_parallel_gzip_levels_and_pairs(tmp_path, parallel_gz, level):
    parallel_gz(50000)   # a few DEFLATE blocks per segment: chains of several segments
    a = [bytes(r) for r in synth.config_short(33, 16000, telomeric=0.01, n_rate=0.001)]
    b = synth.adversarial_short(34, 16000)
    p1, p2 = os.path.join(tmp_path, "a.fastq.gz"), os.path.join(tmp_path, "b.fastq.gz")
    for p, r in ((p1, a), (p2, b)):
        with gzip.open(p, "wb", compresslevel=level) as f:
            f.write(synth.fastq_bytes(r))
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p1, chunk_bytes=50000)
    assert rc == 0 and r1 == a, msg
    rc, msg, r1, r2 = api.ingest_records(api.MODE_PAIR, p1, p2, chunk_bytes=40000)
    assert rc == 0 and r1 == a and r2 == b, msg


@pytest.mark.parametrize("segment", [30000, 4096])   # 4096: no block start inside a segment, the member ends on the sequential decoder
def test_parallel_gzip_errors(tmp_path, parallel_gz, segment):
    import struct
    parallel_gz(segment)
    reads = synth.adversarial_short(35, 12000)
    good = _gz_member(synth.fastq_bytes(reads))
    cases = {
        "truncated body": good[:len(good) // 2],
        "truncated trailer": good[:-3],
        "bad crc": good[:-8] + struct.pack("<I", 12345) + good[-4:],
        "bad isize": good[:-4] + struct.pack("<I", 7),
        "corrupt body": good[:4000] + bytes([good[4000] ^ 0xFF, good[4001] ^ 0x55]) + good[4002:],
        "reserved header flag": good[:3] + b"\x80" + good[4:],
    }
    for name, blob in cases.items():
        p = write(tmp_path, "bad.fastq.gz", blob)
        rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p)
        assert rc != 0, name
        assert "File-IO Error" in msg, (name, msg)


@pytest.mark.parametrize("chunk", [5000, 0])
def test_parallel_gzip_many_small_members(tmp_path, parallel_gz, chunk):
    """A file of many small members (a member is less than a segment): after two of them the members are decoded
    sequentially from their start, and a large member later brings the parallel decoder back -- same records throughout."""
    parallel_gz(40000)
    reads = synth.adversarial_short(41, 1500) + [bytes(r) for r in synth.config_short(42, 14000, telomeric=0.02, n_rate=0.002)]
    data = synth.fastq_bytes(reads)
    cuts = list(range(0, 60000, 3000)) + [60000, len(data) - 9000, len(data) - 6000, len(data) - 3000, len(data)]   # 20 small, one large, 3 small
    blob = b"".join(_gz_member(data[a:b]) for a, b in zip(cuts, cuts[1:]))
    p = write(tmp_path, "many.fastq.gz", blob)
    rc, msg, r1, _ = api.ingest_records(api.MODE_SHORT, p, chunk_bytes=chunk)
    assert rc == 0, msg
    assert r1 == reads

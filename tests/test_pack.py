"""Host packer: ASCII -> planar 2-bit (codes[], src/kmer.cpp:14-31).  CPU only."""
import numpy as np
import pytest

from trew_b200 import api, synth

CODE = {ord("T"): 0, ord("G"): 1, ord("C"): 2, ord("A"): 3}


def unpack(pb):
    off, hi, lo, val = pb.planes()
    out = []
    for r in range(pb.n_reads):
        s = []
        for j in range(off[r], off[r + 1]):
            w, b = divmod(int(j), 32)
            h, l, v = (int(hi[w]) >> b) & 1, (int(lo[w]) >> b) & 1, (int(val[w]) >> b) & 1
            s.append((h << 1 | l) if v else -1)
        out.append(s)
    return out


def expect(read):
    return [CODE.get(c & ~0x20 if chr(c).isalpha() else c, -1) for c in read]


@pytest.mark.parametrize("seed", [1, 2])
def test_pack_matches_codes_table(seed):
    reads = synth.adversarial_short(seed, 300)
    reads += [b"", b"N", b"acgtACGTnN\r", bytes(range(33, 127)), b"A" * 31, b"C" * 32, b"G" * 33, b"T" * 64, b"TTAGGG" * 170]
    buf, locs = api.make_chunk(reads)
    pb = api.PackedBatch(buf, locs)
    got = unpack(pb)
    assert got == [expect(r) for r in reads]
    off = pb.planes()[0]
    assert off[0] == 0 and list(np.diff(off.astype(np.int64))) == [len(r) for r in reads]


def test_pack_invalid_bases_are_zero_coded():
    buf, locs = api.make_chunk([b"NNNNACGT" * 20])
    _, hi, lo, val = api.PackedBatch(buf, locs).planes()
    assert int(np.bitwise_and(hi, ~val).sum()) == 0 and int(np.bitwise_and(lo, ~val).sum()) == 0


def test_pack_matrix_chunk_equals_list_chunk():
    mat = synth.config_short(3, 500)
    b1, l1 = api.matrix_chunk(mat)
    b2, l2 = api.make_chunk([bytes(r) for r in mat])
    assert bytes(b1) == bytes(b2) and (l1 == l2).all()


@pytest.mark.parametrize("simd", ["0", "1", "2"])
@pytest.mark.parametrize("n_ranges,n_threads", [(1, 1), (7, 1), (64, 4), (5000, 3)])
def test_pack_ranges_equal_single_pass_and_list_the_invalid_bases(simd, n_ranges, n_threads):
    """The concurrent range packer (what a staging slot is filled with) produces the same planes as the
    single-threaded one for every cut of the reads, and its list is exactly the zero bits of the val plane."""
    import subprocess, sys, textwrap
    # simd_level() is latched on first use, so each instruction set runs in its own interpreter
    code = textwrap.dedent("""
        import numpy as np
        from trew_b200 import api, synth
        rng = np.random.default_rng(5)
        reads = synth.adversarial_short(3, 400)
        reads += [b"", b"N", b"", b"acgtnN\\r", b"A" * 31, b"N" * 64, b"C" * 65, b"T" * 128, b"TTAGGN" * 170, b""]
        for _ in range(300):
            L = int(rng.integers(0, 400))
            r = rng.choice(np.frombuffer(b"ACGTNacgtn.", dtype=np.uint8), size=L, p=[.22, .22, .22, .22, .04, .02, .02, .01, .01, .01, .01])
            reads.append(bytes(r))
        buf, locs = api.make_chunk(reads)
        ref = api.PackedBatch(buf, locs)
        got = api.PackedBatch(buf, locs, n_ranges=%d, n_threads=%d, want_invalid=True)
        for a, b in zip(ref.planes(), got.planes()):
            assert (a == b).all()
        val = ref.planes()[3]
        bits = np.unpackbits(val.view(np.uint8), bitorder="little")[:ref.bases]
        want = np.flatnonzero(bits == 0)
        assert sorted(got.invalid.tolist()) == want.tolist(), (len(got.invalid), len(want))
        # TREW_PACK_NO_VAL (what the streaming path asks for): same offsets, code planes and list; val unspecified
        lean = api.PackedBatch(buf, locs, n_ranges=%d, n_threads=%d, want_invalid=True, no_val=True)
        for a, b in zip(ref.planes()[:3], lean.planes()[:3]):
            assert (a == b).all()
        assert sorted(lean.invalid.tolist()) == want.tolist()
        print("ok")
    """ % (n_ranges, n_threads, n_ranges, n_threads))
    import os
    env = dict(os.environ, TREW_PACK_SIMD=simd)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("simd", ["0", "1", "2"])
def test_pack_paired_chunk_equals_interleaved_reads(simd):
    """A paired chunk (two buffers, mates alternating in the batch) packs to exactly what the interleaved read list
    packs to -- planes, offsets and invalid-base list -- for every instruction set and cut into ranges."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import numpy as np
        from trew_b200 import api, synth
        a = synth.adversarial_short(8, 700) + [b"", b"N" * 70, b"ACGT" * 40, b"T"]
        b = synth.adversarial_short(9, 700) + [b"ACGTN", b"", b"n" * 129, b"G" * 64]
        inter = [r for pair in zip(a, b) for r in pair]
        ref = api.PackedBatch(*api.make_chunk(inter))
        b1, l1 = api.make_chunk(a)
        b2, l2 = api.make_chunk(b)
        val = ref.planes()[3]
        want = np.flatnonzero(np.unpackbits(val.view(np.uint8), bitorder="little")[:ref.bases] == 0).tolist()
        for n_ranges, n_threads in ((1, 1), (5, 2), (64, 4), (2000, 3)):
            for no_val in (False, True):
                got = api.PackedBatch(b1, l1, n_ranges=n_ranges, n_threads=n_threads, want_invalid=True, no_val=no_val, buf2=b2, locs2=l2)
                assert got.n_reads == ref.n_reads and got.bases == ref.bases
                for x, y in list(zip(ref.planes(), got.planes()))[:3 if no_val else 4]:
                    assert (x == y).all(), (n_ranges, no_val)
                assert sorted(got.invalid.tolist()) == want
        print("ok")
    """)
    env = dict(os.environ, TREW_PACK_SIMD=simd)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]

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

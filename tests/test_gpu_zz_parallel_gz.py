"""The parallel gzip path (pinflate.cpp) through trew_dev_process_file on the GPU box: forced onto small files (no size
threshold, segments of a few DEFLATE blocks) it must give the same tables as submitting the reads directly -- single and
paired, one member and several.  (Last in file order on purpose: everything else has run by then.)"""
import gzip
import os

import pytest

from trew_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_process_file_parallel_gzip(tmp_path, monkeypatch):
    monkeypatch.setenv("TREW_PGZ_MIN_BYTES", "0")
    monkeypatch.setenv("TREW_PGZ_SEGMENT", "40000")
    r1 = synth.adversarial_short(52, 3000, lengths=[100, 150, 151]) + [bytes(r) for r in synth.config_short(53, 12000, telomeric=0.02, n_rate=0.002)]
    r2 = synth.adversarial_short(54, 3000, lengths=[100, 150, 151]) + [bytes(r) for r in synth.config_short(55, 12000, telomeric=0.02, n_rate=0.002)]
    d = str(tmp_path)
    paths = {}
    for tag, reads in (("1", r1), ("2", r2)):
        data = synth.fastq_bytes(reads)
        gz = os.path.join(d, "r%s.fastq.gz" % tag)
        with gzip.open(gz, "wb", compresslevel=6 if tag == "1" else 1) as f:
            f.write(data)
        cut = data.index(b"\n@", len(data) // 3) + 1
        multi = os.path.join(d, "m%s.fastq.gz" % tag)
        open(multi, "wb").write(gzip.compress(data[:cut]) + gzip.compress(data[cut:], 1))
        paths[tag] = {"gz": gz, "multi": multi}
    with api.DeviceContext(api.MODE_SHORT, 5, 32) as ctx:
        ctx.submit_reads(r1)
        want = ctx.finish()
        assert len(want) > 100
        for kind, p in paths["1"].items():
            ctx.reset()
            ctx.process_file(p)
            assert ctx.finish() == want, kind
    with api.DeviceContext(api.MODE_PAIR, 5, 32) as ctx:
        ctx.submit_reads(r1, r2)
        want = ctx.finish()
        for kind in ("gz", "multi"):
            ctx.reset()
            ctx.process_file(paths["1"][kind], paths["2"][kind])
            assert ctx.finish() == want, kind

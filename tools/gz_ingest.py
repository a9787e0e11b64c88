"""File ingest figures (plain / gzip level 1 and 6 / BGZF) through trew_multi_process_file, with TREW_PGZ_TRACE lines."""
import os, sys, time, tempfile, shutil, gzip
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trew_b200 import api, synth
n = 2_000_000
tmp = tempfile.mkdtemp(prefix="trew_gz_")
try:
    plain = os.path.join(tmp, "r.fastq")
    with open(plain, "wb") as f:
        for i in range(0, n, 250_000):
            f.write(synth.fastq_matrix_bytes(synth.config_short(31 + i, 250_000, 150, telomeric=0.01, half_telomeric=0.002, n_rate=0.001, sub=0.01)))
    files = {} if "gz" in sys.argv[1:] else {"plain": plain}
    for lvl in (1, 6):
        p = plain + ".l%d.gz" % lvl
        with open(plain, "rb") as src, gzip.open(p, "wb", compresslevel=lvl) as dst:
            shutil.copyfileobj(src, dst, 1 << 24)
        files["gzip -%d" % lvl] = p
    with api.MultiContext(api.MODE_SHORT, 5, 32, devices=[0]) as m:
        m.set_report_filter(10)
        for name, path in files.items():
            best = None
            for rep in range(3):
                m.reset()
                t0 = time.perf_counter()
                m.process_file(path)
                m.finish_view()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            print("%-10s %7.1f MB: %.3f s = %.2f Gbases/s, %.0f MB/s of FASTQ" % (name, os.path.getsize(path) / 1e6, best, n * 150 / best / 1e9,
                                                                          os.path.getsize(plain) / best / 1e6), file=sys.stderr)
finally:
    shutil.rmtree(tmp, ignore_errors=True)

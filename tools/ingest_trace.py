"""Where a plain-FASTQ file's time goes: TREW_INGEST_TRACE lines of trew_multi_process_file on a 2 M-read file."""
import os, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TREW_INGEST_TRACE"] = "1"
from trew_b200 import api, synth
n = 2_000_000
tmp = tempfile.mkdtemp(prefix="trew_trace_")
try:
    plain = os.path.join(tmp, "r.fastq")
    with open(plain, "wb") as f:
        for i in range(0, n, 250_000):
            f.write(synth.fastq_matrix_bytes(synth.config_short(31 + i, 250_000, 150, telomeric=0.01, half_telomeric=0.002, n_rate=0.001, sub=0.01)))
    with api.MultiContext(api.MODE_SHORT, 5, 32, devices=[0]) as m:
        m.set_report_filter(10)
        for rep in range(3):
            m.reset()
            t0 = time.perf_counter()
            m.process_file(plain)
            t1 = time.perf_counter()
            m.finish_view()
            t2 = time.perf_counter()
            print("run %d: process_file %.2f ms, finish %.2f ms -> %.1f Gbases/s" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, n * 150 / (t2 - t0) / 1e9), file=sys.stderr)
finally:
    shutil.rmtree(tmp, ignore_errors=True)

"""Wall clock of the `trew` binary's phases on a small and a 2 M-read file (TREW_CLI_TIMING)."""
import os, sys, time, tempfile, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trew_b200 import api, synth
tmp = tempfile.mkdtemp(prefix="trew_cli_")
try:
    for n in (20_000, 2_000_000):
        p = os.path.join(tmp, "r%d.fastq" % n)
        with open(p, "wb") as f:
            for i in range(0, n, 250_000):
                f.write(synth.fastq_matrix_bytes(synth.config_short(31 + i, min(250_000, n - i), 150, telomeric=0.01, half_telomeric=0.002, n_rate=0.001, sub=0.01)))
        for devs in ("", "all"):
            env = dict(os.environ, TREW_CLI_TIMING="1")
            if devs:
                env["TREW_DEVICES"] = devs
            for rep in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([api.CLI_PATH, "short", "5", "32", p], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
                print("n=%d TREW_DEVICES=%s run %d wall %.3f s\n%s" % (n, devs or "(default)", rep, time.perf_counter() - t0, r.stderr.decode()))
finally:
    shutil.rmtree(tmp, ignore_errors=True)

"""Driver of tests/test_exact_thread.py::test_thread_path_under_sanitizers: runs the host build of the thread-per-read
exact routing (short, paired, long; edge lengths, N runs, truncated mates) inside a library compiled with
-fsanitize=address,undefined.  compute-sanitizer is closed on the GPU pool, so this is the memory / shift check of the
code the thread kernels run.  usage: LD_PRELOAD=libasan.so python exact_thread_sanitize.py libetc_asan.so"""
import os
import sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_exact_thread as t
from trew_b200 import synth
lib = C.CDLL(sys.argv[1])
lib.etc_scan_reads.restype = C.c_long
lib.etc_scan_reads.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
lib.etc_scan_pairs.restype = C.c_long
lib.etc_scan_pairs.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
lib.etc_scan_long.restype = C.c_long
lib.etc_scan_long.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p]
for mn, mx, lengths in [(5,32,[150,160,128,100,75,36,12,9,6]), (3,32,[160,159,97,7]), (5,8,[33,40,150])]:
    reads = synth.adversarial_short(5+mx, 1500, max_unit=mx, lengths=lengths) + [b"", b"A", b"N"*160, b"A"*160, b"TTAGGG"*26+b"TTAG"]
    t.run(lib, reads, mn, mx)
r1, r2 = synth.adversarial_pairs(9, 600, read_len=150, max_unit=32, truncate_mate2=0.2)
t.run_pairs(lib, r1, r2, 5, 32)
reads = synth.adversarial_long(11, 60, min_len=150, max_len=4000, max_unit=32)
t.run_long(lib, reads, 5, 32, 150); t.run_long(lib, reads, 5, 32, 160)
print("sanitizer run clean")

"""The C-ABI library loads and exports every symbol include/trew_b200.h declares.  No compute, no GPU."""
import ctypes
import os
import re

from trew_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "trew_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(trew_[a-z0-9_]+)\s*\(", text))
    names.discard("trew_chunk_sink")
    return sorted(names)


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(api.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.trew_abi_version() == 2


def test_status_strings():
    lib = api.load_library()
    assert lib.trew_status_string(0) == b"ok"
    # the reference's own message for an over-long short read (src/kmer.cpp:1007)
    assert lib.trew_status_string(4) == b"This mode is designed for short-read sequencing. Please use 'trew long'."


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "trew_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle", "").replace("The oracle", "") or f == "api.py", f

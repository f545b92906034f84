"""CPU: the C-ABI shared library loads and exports every symbol include/touchnet_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from touchnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "touchnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/touchnet_b200.h but not exported"


def test_python_binding_covers_the_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == _header_symbols()
    lib = _lib.load()
    assert lib.tn_version() == 100


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The evidence the profiling guide asks for: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA) in the SASS."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic


def test_ctypes_signatures_match_the_header_arity():
    """Every prototype of include/touchnet_b200.h has a ctypes signature in touchnet_b200/_lib.py with the same number
    of parameters (a drift here corrupts the call stack silently)."""
    import os
    import re
    from touchnet_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "touchnet_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|const char\*|int64_t)\s+(tn_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S)
    assert len(protos) >= 30
    for name, args in protos:
        if name == "tn_last_error":
            continue
        n = 0 if args.strip() in ("void", "") else len(args.split(","))
        assert name in _lib._SIGNATURES, f"{name} is declared in the header but has no ctypes signature"
        assert len(_lib._SIGNATURES[name]) == n, (name, len(_lib._SIGNATURES[name]), n)
    assert set(_lib._SIGNATURES) <= {p[0] for p in protos}, "ctypes signature without a header prototype"

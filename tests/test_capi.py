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


def test_argument_errors_of_the_entry_points_added_late_in_round_1():
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU: non-zero return code +
    tn_last_error() -> TouchNetB200Error (the reference's convention is Python exceptions)."""
    import ctypes
    import pytest
    from touchnet_b200 import _lib
    arr = (ctypes.c_void_p * 2)(16, 32)
    bad_calls = [
        ("tn_peer_reduce_scatter_f32", (arr, 0, 0, 64, 4, 1.0, 32, None), "n_peers"),
        ("tn_peer_reduce_scatter_f32", (arr, 2, 3, 64, 4, 1.0, 32, None), "multiple of 4"),
        ("tn_peer_all_gather", (arr, 2, 24, 64, 32, None), "multiple of 16"),
        ("tn_set_gemm_group", (0,), "out of range"),
        ("tn_set_sm_margin", (-1,), "out of range"),
        ("tn_adamw_f32", (None, None, None, None, None, 4, 1e-3, 0.9, 0.95, 1e-8, 0.1, 0.1, 0.1, None, None), "null"),
        ("tn_pack_layout_i64", (None, None, None, None, None, None, 0, 1, 8, 0, 1, 2, None, None, None, None, None, None), "null"),
        ("tn_sumsq_f32", (None, 4, None, None, None), "null"),
    ]
    for name, args, needle in bad_calls:
        with pytest.raises(_lib.TouchNetB200Error, match=needle):
            _lib.call(name, *args)
    _lib.call("tn_set_gemm_group", 8)          # valid values are accepted without a device
    _lib.call("tn_set_sm_margin", 0)

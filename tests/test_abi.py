"""CPU: the C-ABI library loads and exports every symbol include/hbr.h declares; the ctypes table matches."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hbr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hbr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from human_body_reconstruction_b200 import _lib
    so = _lib.build()
    l = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(l, s), f"{s} declared in hbr.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert l.hbr_abi_version() == 2
    assert not any(s.startswith("hbr_debug") for s in syms), "probe entry points belong to the debug library"


def test_param_count_and_errors_without_gpu():
    from human_body_reconstruction_b200 import _lib
    l = _lib.lib()
    assert l.hbr_mlp_param_count(ctypes.byref(_lib.MlpDims(32, 24))) == 14227
    assert l.hbr_mlp_param_count(ctypes.byref(_lib.MlpDims(100, 24))) == -1
    assert b"in0" in l.hbr_last_error()
    g = _lib.HashGeom()
    g.L, g.F, g.T = 0, 2, 16                      # invalid geometry is rejected before any CUDA call
    assert l.hbr_hash_encode_fwd(None, 0, 4, None, ctypes.byref(g), None, 0, None) == -1


def test_sass_is_sm100a_only():
    import subprocess
    from human_body_reconstruction_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_debug_library_is_separate():
    """The probes / self tests (csrc/debug) build into their own library; the product library exports none of them."""
    from human_body_reconstruction_b200 import _lib
    prod = ctypes.CDLL(_lib.build())
    dbg = ctypes.CDLL(_lib.build_debug())
    for name in _lib.DEBUG_SIGNATURES:
        assert hasattr(dbg, name)
        assert not hasattr(prod, name), f"{name} leaked into the product ABI"

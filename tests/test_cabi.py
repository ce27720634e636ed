"""CPU: the C-ABI library loads and exports every symbol include/ngp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import _ngp_b200 as nb
from util import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[a-zA-Z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(nb.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ngp_b200.h but not exported"


def test_binding_table_matches_header():
    assert set(_declared()) <= set(nb.EXPORTED) | {"ngp_stream_t"}
    lib = nb.load()
    assert lib.ngp_version() == 1
    assert lib.ngp_build_arch() == b"sm_100a"


def test_sass_is_blackwell_native():
    """tcgen05 / TMEM show up as UTCHMMA / LDTM in the SASS of the shipped library (B200_PROFILING.md)."""
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        import pytest; pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", nb.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync / wmma path
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", nb.LIB_PATH], capture_output=True, text=True).stdout


def _prototypes():
    """name -> list of (type, name) parameters, parsed from include/ngp_b200.h"""
    hdr = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|void|const char\*)\s+(ngp_[a-zA-Z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), " ".join(m.group(2).split())
        protos[name] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return protos


def test_ctypes_signatures_match_the_header():
    """Every stream-taking entry point bound in _ngp_b200._SIGNATURES has exactly the header's parameter list: same count, a trailing
    ngp_stream_t, pointers bound as c_void_p, floats as c_float, 64-bit integers as c_uint64 / c_size_t (an ABI mismatch of this kind only
    shows up on the GPU otherwise)."""
    protos = _prototypes()
    assert len(protos) >= 60
    c = ctypes
    for name, argtypes in nb._SIGNATURES.items():
        assert name in protos, f"{name} is bound but not declared"
        params = protos[name]
        if name in ("ngp_ffmlp_allocate_splitk", "ngp_ffmlp_free_splitk"):
            assert len(params) == len(argtypes)
            continue
        assert params and params[-1].startswith("ngp_stream_t"), f"{name}: last parameter must be the stream"
        assert len(params) - 1 == len(argtypes) - 1 or len(params) == len(argtypes), (name, len(params), len(argtypes))
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, binding {len(argtypes)}"
        for p, t in zip(params, argtypes):
            if "*" in p or p.startswith("ngp_stream_t"):
                assert t is c.c_void_p, (name, p, t)
            elif p.startswith("float"):
                assert t is c.c_float, (name, p, t)
            elif p.startswith("uint64_t"):
                assert t is c.c_uint64, (name, p, t)
            elif p.startswith("size_t"):
                assert t is c.c_size_t, (name, p, t)
            elif p.startswith("uint32_t"):
                assert t is c.c_uint32, (name, p, t)
            elif p.startswith("int"):
                assert t is c.c_int, (name, p, t)
            else:
                raise AssertionError(f"{name}: unhandled parameter type in the header: {p}")

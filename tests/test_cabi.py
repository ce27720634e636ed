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

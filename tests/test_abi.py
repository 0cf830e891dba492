"""The C-ABI library loads and exports every symbol include/ccd_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT
from collisiondetection_b200 import api


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ccd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ccd_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    lib = api.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libccd_b200.so does not export %s" % n
    assert set(names) == set(api.EXPORTS)


def test_version_string():
    lib = api.load_library()
    assert b"sm_100a" in lib.ccd_version()


def test_create_without_device_fails_loudly():
    """On a box without a GPU ccd_create must fail (CCD_ERR_NODEVICE) — there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        return
    lib = api.load_library()
    h = ctypes.c_void_p()
    rc = lib.ccd_create(ctypes.byref(h), 0)
    assert rc == -4 and not h
    try:
        api.Context(0)
    except api.CcdError:
        pass
    else:
        raise AssertionError("Context() succeeded without a CUDA device")


def test_product_does_not_import_oracle():
    """The product path must not route through the CPU checkers."""
    pkg = os.path.join(ROOT, "collisiondetection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/ccd_oracle.c: orc_roots01", ""), f

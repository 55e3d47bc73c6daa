"""CPU-side checks of the drop-in boundary: the CUDA library builds for sm_100a, loads, exports every
symbol include/gmg_b200.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from helpers import pkg


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gmg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmg_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return pkg().capi.load_library()


def test_every_declared_symbol_is_exported_and_bound(lib):
    capi = pkg().capi
    names = header_symbols()
    assert len(names) >= 35
    assert sorted(capi.SIGNATURES) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", pkg().capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.gmg_create(0, ctypes.byref(h)) == -2  # GMG_ENODEVICE
    with pytest.raises(pkg().capi.GmgError):
        pkg().capi.Gmg(0)

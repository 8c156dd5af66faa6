"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports every symbol the
header declares; the product never routes through the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "task-level-insights-from-eigenvalues-across-sequence-models_b200")


@pytest.fixture(scope="module")
def lib():
    import eigb200.build as b
    b.build()
    import eigb200._lib as L
    return L.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "eigb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eigb200_[a-z0-9_]+)\s*\(", src)))


def test_every_header_symbol_is_exported_and_bound(lib):
    import eigb200._lib as L
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
        assert s in L.SIGNATURES, "ctypes binding lacks %s" % s
    assert sorted(L.SIGNATURES) == syms


def test_version_and_error_plumbing(lib):
    assert lib.eigb200_version() == 100
    rc = lib.eigb200_mamba2_eig(None, None, 0, 1, 1, 4, None, None, None, 1, None, 1, None, None, 0, 0, None, 0.0)
    assert rc == -1 and b"null" in lib.eigb200_last_error()


def test_no_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    sm = C.c_int(0)
    assert lib.eigb200_device_info(0, C.byref(sm), None, None) == -2          # EIGB200_ECUDA, loudly
    import eigb200.ops as ops
    import eigb200
    with pytest.raises(eigb200.Eigb200Error):
        ops.mamba2_eig(torch.zeros(1, 4, 8), torch.zeros(1, 8), torch.zeros(1), torch.zeros(1))


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_built_for_sm100a_only(lib):
    import subprocess, shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", os.path.join(PKG, "libeigb200.so")], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs

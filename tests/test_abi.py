"""The C-ABI library loads and exports every symbol include/alchemy_b200.h declares; without a CUDA
device every compute entry fails loudly with a status code (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "alchemy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_and_binding_agree():
    from alchemy_b200 import _lib
    names = header_functions()
    assert len(names) >= 55
    assert sorted(_lib.SIGNATURES) == names


def test_generated_haskell_binding_matches_the_abi():
    """haskell/.../B200/FFI.hs (the `foreign import`s a maintainer of the reference would add) is generated from the ABI table and
    committed: it must be exactly what the generator prints, and name every entry of the header"""
    import subprocess
    import sys
    gen = os.path.join(ROOT, "tools", "gen_haskell_ffi.py")
    assert subprocess.run([sys.executable, gen, "--check"]).returncode == 0, "run tools/gen_haskell_ffi.py"
    hs = open(os.path.join(ROOT, "haskell", "Crypto", "Lol", "Cyclotomic", "Tensor", "B200", "FFI.hs")).read()
    for name in header_functions():
        assert f'ccall unsafe "{name}"' in hs, name


def test_library_exports_every_declared_symbol():
    from alchemy_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = C.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(handle, name), name
    lib = _lib.lib()
    assert lib.ab_version().decode().startswith("alchemy_b200")


def test_product_never_imports_the_oracle():
    """nothing under alchemy_b200/ may import, include, link or dlopen anything under oracle/"""
    pkg = os.path.join(ROOT, "alchemy_b200")
    bad = re.compile(r"^\s*(from|import)\s+\.*oracle|import_module\(.*oracle|#\s*include.*oracle|libalchemy_oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{f} reaches into the oracle"


def test_no_device_means_error_not_fallback():
    import torch

    from alchemy_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _lib.lib()
    assert lib.ab_device_count() == 0
    pe, npe = _lib.prime_exps(64)
    qs = (C.c_int64 * 1)(193)
    h = C.c_void_p()
    rc = lib.ab_plan_create(C.byref(h), pe, npe, qs, 1, None)
    assert rc == _lib.AB_ERR_CUDA and b"no CPU fallback" in lib.ab_last_error()
    # the lol-cpp-shaped entry reports the same failure instead of computing on the host
    import numpy as np
    y = np.arange(32, dtype=np.int64)
    keep = y.copy()
    rc = lib.tensorLRq(1, y.ctypes.data_as(C.c_void_p), 64, pe, npe, qs)
    assert rc == _lib.AB_ERR_CUDA and (y == keep).all()
    assert lib.ab_plan_n(None) == 0 and lib.ab_gadget_len(None, 1, 0) == -1
    lib.ab_plan_destroy(None)

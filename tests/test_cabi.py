"""The C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/ecb200.h
declares, and refuses to work (loudly) when no B200 is present.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_symbols(header="ecb200.h", prefix="ecb_"):
    with open(os.path.join(ROOT, "include", header)) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, text)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for must in ("ecb_create", "ecb_push", "ecb_finalize", "ecb_destroy", "ecb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    from alntools_b200 import _native
    lib = _native.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), "libecb200.so does not export %s" % name
        assert name in _native.SIGNATURES, "no ctypes signature for %s" % name
    assert lib.ecb_version() >= 1000


def test_host_emitter_library_exports_every_declared_symbol(built):
    from alntools_b200 import bamcols
    lib = bamcols.load_library()
    syms = _declared_symbols("bamcols.h", "bamcols_")
    for must in ("bamcols_open", "bamcols_emit", "bamcols_set_tables", "bamcols_close"):
        assert must in syms
    for name in syms:
        assert hasattr(lib, name), "libbamcols.so does not export %s" % name
        assert name in bamcols.SIGNATURES, "no ctypes signature for %s" % name


def test_library_contains_sm100a_code(built):
    from alntools_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from alntools_b200._native import EcBuilder, EcbError
    with pytest.raises(EcbError) as info:
        EcBuilder(10, 2)
    assert info.value.code == -3 and "no CPU path" in str(info.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "alntools_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as fh:
                    text = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "/root/reference" not in text, fn

"""CPU-side checks of the boundary: the CUDA product library loads and exports every symbol include/vcb200.h declares
(no compute calls without a GPU), and refuses to create a context when no CUDA device exists (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from visioncortex_b200._ffi import DEFAULT_LIB, EXPORTED_SYMBOLS, Library, VcbError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "visioncortex_b200", "csrc"), "-s"])
    return Library(DEFAULT_LIB)


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_symbol_list_agree():
    assert header_symbols() == sorted(EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(product_lib):
    for s in header_symbols():
        assert hasattr(product_lib.lib, s), s


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device, context creation must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from visioncortex_b200.runtime import Context

    with pytest.raises(VcbError):
        Context(0, lib=product_lib)


def test_sass_is_sm100a(product_lib):
    out = subprocess.run(["cuobjdump", "-lelf", DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out

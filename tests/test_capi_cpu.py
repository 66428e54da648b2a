"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/ppp_b200.h declares
(no compute calls without a GPU), and fails loudly -- never silently falls back -- when there is no device."""
import ctypes
import os
import re

import pytest

from prophylo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ppp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ppp_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(capi.EXPORTS)
    assert lib.ppp_abi_version() == 1
    assert lib.ppp_words_per_row(1024) == 32 and lib.ppp_words_per_row(1459) == 48 and lib.ppp_words_per_row(1) == 4


def test_hit_layout_is_32_bytes():
    assert capi.HIT_DTYPE.itemsize == 32
    assert capi.HIT_DTYPE.fields["score"][1] == 8 and capi.HIT_DTYPE.fields["margin"][1] == 28
    assert ctypes.sizeof(capi.Filter) == 16


def test_no_silent_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PPPError) as ei:
        capi.Context(0)
    assert ei.value.code == -2 and "no CPU path" in str(ei.value)

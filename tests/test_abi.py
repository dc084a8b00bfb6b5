"""CPU: the C-ABI library loads, exports every symbol include/fcsearch.h declares, and reports errors as
codes (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from merizo_search_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for header in ("fcsearch.h", "fcsembed.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(fcs_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported():
    lib = native.load()
    names = _declared()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), f"{name} declared in fcsearch.h but not exported by libfcsearch.so"
    assert sorted(native.EXPORTS) == names, "native.EXPORTS and include/*.h disagree"


def test_version_and_error_string():
    lib = native.load()
    assert lib.fcs_version() >= 100
    assert isinstance(lib.fcs_last_error(), bytes)


def test_invalid_arguments_are_codes():
    lib = native.load()
    h = C.c_void_p()
    assert lib.fcs_db_create(0, 10, 64, 0, 0, C.byref(h)) == native.ERR_INVALID      # dim != 128
    assert b"dim" in lib.fcs_last_error()
    assert lib.fcs_db_create(0, 0, 128, 0, 0, C.byref(h)) == native.ERR_INVALID       # no rows
    assert lib.fcs_db_create(0, 10, 128, 2**32, 0, C.byref(h)) == native.ERR_INVALID  # ids beyond u32
    assert lib.fcs_db_create(0, 10, 128, 0, 0x80, C.byref(h)) == native.ERR_INVALID   # unknown flag
    assert lib.fcs_db_finalize(None) == native.ERR_INVALID
    assert lib.fcs_search(None, None, 1, None, 0.0, 1, 0, 0, 0, None, None) == native.ERR_INVALID
    assert lib.fcs_db_destroy(None) == native.OK
    e = C.c_void_p()
    assert lib.fcs_embedder_create(0, None, 2, None, 3000, C.byref(e)) == native.ERR_INVALID
    assert lib.fcs_embed(None, None, None, 1, None) == native.ERR_INVALID
    assert b"null embedder" in lib.fcs_last_error()
    assert lib.fcs_embedder_destroy(None) == native.OK


def test_no_gpu_fails_loudly_not_silently():
    """Without a GPU the product path must raise -- there is no CPU fallback behind the binding."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.FcsError):
        native.Database(100)
    with pytest.raises(native.FcsError):
        native.device_count()


def test_struct_layouts_match_header():
    assert C.sizeof(native.Timing) == 24
    assert C.sizeof(native.Info) == 48
    assert C.sizeof(native.EgnnWeights) == 80
    assert C.sizeof(native.EmbedTiming) == 32

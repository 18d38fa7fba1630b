"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/balm_b200.h
declares, and fails loudly (no silent CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from balm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "balm_b200.h")).read()
    declared = set(re.findall(r"\b(balm_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"balm_ctx"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.balm_version() >= 100


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import balm_b200
    with pytest.raises(balm_b200.BalmError) as e:
        balm_b200.Context(4)
    assert e.value.status == 2  # BALM_ERR_CUDA, never a CPU fallback


def test_struct_layouts_match_header():
    from balm_b200 import _lib
    assert C.sizeof(_lib.LmOpts) == 56 and C.sizeof(_lib.Trace) == 64 and C.sizeof(_lib.Timings) == 72


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "balm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test oracle", ""), f

"""CPU: the C-ABI library builds, loads and exports every symbol include/lsm_b200.h declares.
No compute calls here (no GPU in this suite)."""
import os
import re

import pytest

from lsm_speech_classifier_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lsm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    so = build.build()
    assert os.path.exists(so)
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lsm_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.LsmError):
        _lib.Context(0)
    with pytest.raises(_lib.LsmError):
        _lib.context()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lsm_speech_classifier_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liblsm_oracle" not in src, f


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.FrontendParams) == 10 * 4 + 16 * 8
    assert C.sizeof(_lib.ReservoirParams) == 6 * 4 + 8


def test_host_design_helpers_agree_with_numpy():
    """lsm_gammatone_design / lsm_zoom_table are plain host code (no device work): check them here."""
    import ctypes as C
    import numpy as np
    from lsm_speech_classifier_b200 import filterbank as fb
    lib = _lib.load()
    for ch in (40, 64, 128, 256):
        out = np.zeros((ch, 10))
        assert lib.lsm_gammatone_design(16000.0, ch, 50.0, C.c_void_p(out.ctypes.data)) == 0
        np.testing.assert_allclose(out, fb.gammatone_coefs(16000, ch, 50), rtol=1e-12, atol=0)
    for n_in in (98, 101):
        i0, f = np.zeros(100, np.int32), np.zeros(100)
        assert lib.lsm_zoom_table(n_in, 100, C.c_void_p(i0.ctypes.data), C.c_void_p(f.ctypes.data)) == 0
        a, b = fb.zoom_table(n_in, 100)
        assert np.array_equal(i0, a) and np.array_equal(f, b)
    assert lib.lsm_gammatone_design(16000.0, 0, 50.0, None) != 0

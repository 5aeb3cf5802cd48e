"""CPU: the C-ABI library loads and exports every symbol include/apg_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "apg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from ap_gym_b200 import _capi

    if not os.path.exists(_capi.LIB_PATH):
        from ap_gym_b200._build import build

        build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/apg_b200.h but not exported"
        assert name in _capi.SYMBOLS, f"{name} has no ctypes prototype in ap_gym_b200/_capi.py"
    assert _capi.lib().apg_version() == 100


def test_ctypes_struct_layouts_match_the_compiled_header():
    from ap_gym_b200 import _capi

    out = (ctypes.c_int32 * 5)()
    _capi.lib().apg_struct_sizes(out)
    mine = [ctypes.sizeof(c) for c in (_capi.ImageCfg, _capi.ImageBuffers, _capi.LidarCfg, _capi.LidarBuffers,
                                       _capi.Pcg64State)]
    assert list(out) == mine, f"library {list(out)} vs ctypes {mine}"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    import ap_gym_b200 as apg

    ds = apg.ArrayImageDataset(np.zeros((4, 28, 28, 1), np.uint8), np.zeros(4, np.int32), 2)
    with pytest.raises(Exception) as ei:
        apg.ImageClassificationVectorEnv(2, apg.ImagePerceptionConfig(dataset=ds))
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "active-perception-gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"

"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/carmel_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "carmel_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cml_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native_lib):
    import carmel_b200 as cb
    lib = cb.load_library()
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/carmel_b200.h but not exported"
    assert sorted(cb.exported_symbols()) == decl


def test_no_gpu_means_loud_failure(native_lib):
    import torch
    import carmel_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CarmelB200Error) as ei:
        cb.Context(0, 64, 0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_reference_oracle():
    """the product path may not include, link or execute anything under oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "carmel_b200")):
        if "_build" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"oracle/|carmel_oracle|orc::", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad

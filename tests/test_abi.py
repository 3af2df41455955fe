"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/ppcsr_b200.h declares; without a GPU the product fails loudly (no CPU fallback)."""
import importlib
import os
import re

import pytest

pp = importlib.import_module("parallel-packed-csr_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppcsr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppcsr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = pp.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/ppcsr_b200.h but not exported"
        assert name in pp.SYMBOLS, f"{name} has no ctypes prototype"
    assert set(pp.SYMBOLS) == set(names)


def test_struct_layouts_match_header():
    import ctypes as C

    assert C.sizeof(pp.BatchStats) == 12 * 8 + 2 * 4 + 6 * 4 + 2 * 4
    assert C.sizeof(pp.Geometry) == 32
    assert C.sizeof(pp.InvariantReport) == 80


def test_no_cpu_fallback():
    L = pp.load_library()
    if L.ppcsr_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pp.PpcsrError):
        pp.Shard(10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "parallel-packed-csr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("oracle_py", "pcsr_oracle", "oracle/_ref", "oracle/_build", "ref_driver\""):
                    assert needle not in text, (f, needle)

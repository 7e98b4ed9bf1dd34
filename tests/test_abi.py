"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/adaface_b200.h declares
(no compute calls without a GPU); the product package never imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    import adaface_dev_b200 as a
    lib = ctypes.CDLL(a._lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "adaface_b200.h")).read()
    declared = set(re.findall(r"\b(adaface_[a-z0-9_]+)\s*\(", header))
    assert declared == set(a._lib.SIGNATURES), "ctypes table and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None
    assert a._lib.load().adaface_version() == 3


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "adaface-dev_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    import adaface_dev_b200 as a
    x = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        a.ops.proj(x, x)
    proc, attn = a.AttnProcessor_LoRA_Capture(), a.Attention(320, None, 8, 40)
    with pytest.raises(RuntimeError):
        proc(attn, torch.zeros(1, 4, 320))

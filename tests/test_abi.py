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
    assert a._lib.load().adaface_version() == 4


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


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """The boundary is a C ABI: a C99 translation unit that includes include/adaface_b200.h compiles with -pedantic, links
    against libadaface_b200.so and calls an entry point (no compute: there is no GPU here)."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    g.build()
    import adaface_dev_b200 as a
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    src = tmp_path / "abi_demo.c"
    src.write_text('#include "adaface_b200.h"\n'
                   "int main(void) { return adaface_version() == ADAFACE_B200_ABI_VERSION ? 0 : 1; }\n")
    libdir = os.path.dirname(a._lib.LIB_PATH)
    exe = tmp_path / "abi_demo"
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-ladaface_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert subprocess.run([str(exe)]).returncode == 0

"""The C-ABI library loads and exports every symbol include/zkodst.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zkodst.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", src)))


def test_exports(zk):
    lib = ctypes.CDLL(zk.library_path())
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), s


def test_no_device_is_an_error_not_a_fallback(zk):
    import torch
    if torch.cuda.is_available():
        return
    try:
        zk.Context(0)
    except zk.ZkError as e:
        assert e.code == -2
    else:
        raise AssertionError("context creation must fail without a CUDA device")


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "zk-odst_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libzkoracle" not in text and "oracle/" not in text, f

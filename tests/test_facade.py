"""include/zkodst.hpp — the reference's plugin surface (Blake2fInstructions, Table16Chip, Blake2f gadget,
Params / keygen / create_proof / verify_proof / MockProver) in C++ over the C ABI.  tests/cpp/facade_test.cpp is
the reference's bench and commented test module rewritten against it; here it is built with g++ (CPU) and run
(GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "zk-odst_b200")


def build(tmp_path):
    import zk_odst_b200 as zk
    zk.load_library()  # builds nothing; fails loudly if libzkodst.so is missing
    exe = str(tmp_path / "facade_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-L", LIBDIR, "-lzkodst",
           "-Wl,-rpath," + LIBDIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_facade_builds_and_fails_loudly_without_a_device(tmp_path):
    """Header-only C++17 over the C ABI: compiles warning-free and links against libzkodst.so.  Without a
    CUDA device the first call throws zkodst::Error (Backend, ZK_E_CUDA): there is no CPU fallback."""
    import torch
    exe = build(tmp_path)
    # the gadget, the recording layouter and the circuit's lay-out are host code: they run everywhere
    res = subprocess.run([exe, "--host-only"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "facade host part ok" in res.stdout, res.stdout + res.stderr
    if torch.cuda.is_available():
        pytest.skip("a device is present: test_facade_runs covers the run")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2
    assert "code -2" in res.stderr and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_facade_runs(tmp_path):
    exe = build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "facade ok" in res.stdout

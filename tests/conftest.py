import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def a2ds():
    """the product package (directory name has a hyphen, hence importlib)"""
    return importlib.import_module("a2d-shells_b200")


@pytest.fixture(scope="session")
def orc():
    """plain-C oracle (test infrastructure)"""
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference build, when oracle/_ref is present"""
    import refdrv
    if not refdrv.available():
        pytest.skip("oracle/_ref not built")
    refdrv.lib()
    return refdrv


@pytest.fixture(scope="session")
def emul():
    """host emulation of the kernel's per-lane math (tests/host_emul.cpp)"""
    import ctypes as C
    so = os.path.join(ROOT, "tests", "_host_emul.so")
    src = os.path.join(ROOT, "tests", "host_emul.cpp")
    hdrs = [os.path.join(ROOT, "a2d-shells_b200", "csrc", h) for h in ("mitc4_math.h", "mitc4_tying.h", "mitc9_math.h")]
    if (not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] +
                                                             [os.path.getmtime(h) for h in hdrs])):
        # -mfma + contraction on: mimics nvcc's FMA fusion outside the strict sections
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-mfma", "-ffp-contract=fast", "-fPIC",
                               "-shared", "-o", so, src])
    return C.CDLL(so)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """Build libb200reg.so and the oracle once per session (no GPU needed)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from platipy_b200.engine import Engine

    return Engine.get(0)


@pytest.fixture(scope="session")
def emu():
    """The barrier-free kernels (platipy_b200/csrc/*_kernels.cuh) compiled for the HOST by g++ behind tests/emu/cuda_emu.h, a
    serial emulation of blockIdx / threadIdx: this container has no GPU, so the CPU suite checks those kernels' indexing and
    arithmetic against the oracle this way.  Test infrastructure; the package never loads it."""
    import ctypes
    import glob
    import subprocess

    here = os.path.join(ROOT, "tests", "emu")
    build = os.path.join(here, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libemu_distmap.so")
    srcs = [os.path.join(here, "emu_distmap.cpp"), os.path.join(here, "cuda_emu.h")] + glob.glob(os.path.join(ROOT, "platipy_b200", "csrc", "*_kernels.cuh"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    return ctypes.CDLL(so)

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dropin")):
    if p not in sys.path:
        sys.path.insert(0, p)

BUILD = os.path.join(ROOT, "tests", "_build")
CSRC = os.path.join(ROOT, "curdleproofs_pie_b200", "csrc")
SEAM_SO = os.path.join(BUILD, "libcpg_hostseam.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_seam():
    """TEST SEAM: the kernels' per-thread functors compiled for the host (PTX carry flag emulated,
    bigint.cuh) behind the same C ABI, so the CPU tier can check the launch logic bit-for-bit.
    Never shipped, never loaded by the product package."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "cpg.h")]
    if _newer(SEAM_SO, srcs):
        return SEAM_SO
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(
        ["g++", "-O2", "-fopenmp", "-x", "c++", "-std=c++17", "-DCPG_HOST_EMU", "-shared", "-fPIC", "-pthread",
         "-o", SEAM_SO, os.path.join(CSRC, "cpg_api.cu")]
    )
    return SEAM_SO


@pytest.fixture(scope="session")
def seam_lib():
    from curdleproofs_pie_b200 import runtime

    lib = runtime.CpgLib(build_seam(), 0)
    assert lib.backend == "host-emulation-test-seam"
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    from curdleproofs_pie_b200 import runtime

    lib = runtime.get_lib()  # raises if the CUDA library or the GPU is missing
    assert lib.backend == "cuda-sm_100a", lib.backend
    return lib


@pytest.fixture(scope="session")
def cref():
    from oracle import cref_binding

    return cref_binding.load()

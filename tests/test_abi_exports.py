"""The product library loads and exports every symbol include/cpg.h declares (no compute calls:
this tier has no GPU), and refuses to initialise without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    import __graft_entry__ as ge

    ge.build()
    return ctypes.CDLL(ge.LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cpg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cpg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(product_lib):
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(product_lib, n), n


def test_binding_covers_header():
    from curdleproofs_pie_b200 import runtime

    assert set(runtime.EXPORTS) == set(declared_symbols())


def test_product_library_is_cuda_and_has_no_cpu_fallback(product_lib):
    product_lib.cpg_backend.restype = ctypes.c_char_p
    assert product_lib.cpg_backend() == b"cuda-sm_100a"
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is for CPU-only hosts")
    from curdleproofs_pie_b200 import runtime

    with pytest.raises(runtime.CpgError):
        runtime.CpgLib(runtime.LIB_PATH, 0)
    product_lib.cpg_sync.restype = ctypes.c_int
    assert product_lib.cpg_sync() != 0          # nothing works before a successful cpg_init


def test_sass_is_sm100a_integer_pipe(product_lib):
    """The kernels are sm_100a SASS whose multiply work is IMAD.WIDE/IMAD on the integer pipe."""
    import shutil
    import subprocess

    import __graft_entry__ as ge

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", ge.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out

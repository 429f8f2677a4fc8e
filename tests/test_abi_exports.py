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


def test_sass_is_sm100a_integer_pipe(product_lib, tmp_path):
    """The kernels are sm_100a SASS whose field arithmetic is IMAD.WIDE.U32 on the integer pipe: the Fq product the
    point kernels call is ~276 wide multiply-accumulates (the 300-MAC CIOS minus folded constants) with no memory
    access, the square ~210, and no kernel holds a tensor-core or TMA instruction (this path is not a contraction).
    tools/sass_report.py writes the same figures to profiles/r02_sass_histogram.txt."""
    import json
    import shutil
    import subprocess
    import sys

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_report.py"), str(tmp_path / "t")], capture_output=True, text=True, check=True).stdout
    rep = json.loads(out)
    assert rep["arch"] == ["sm_100a"]
    assert rep["total"]["tensor_or_tma"] == 0
    ks = rep["kernels"]
    assert len(ks) >= 60
    for name in ("BucketAccumulate<128,3>", "Decompress<128,3>", "FixedMsmWindow<128,3>", "WindowReduce<128,3>"):
        assert name in ks, name
    subs = rep["bucket_accumulate_subs"]
    muls = [s_ for s_ in subs if 270 <= s_["wide"] <= 300]
    sqrs = [s_ for s_ in subs if 200 <= s_["wide"] <= 230]
    assert muls and sqrs
    for s_ in muls + sqrs:
        assert s_["mem"] == 0                       # register ABI: operands never touch memory
        assert s_["wide"] / s_["instrs"] > 0.5      # the body is mostly the multiply pipe
    # the two dominant kernels keep their state in registers (no local-memory spills in the loop bodies)
    assert ks["BucketAccumulate<128,3>"]["wide"] >= 480       # its own body: two inlined products beside the calls

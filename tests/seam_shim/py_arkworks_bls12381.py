"""TEST SHIM (CPU tier only): lets the UNMODIFIED reference import `py_arkworks_bls12381` and get the
drop-in surface (dropin/py_arkworks_bls12381) running on the host-emulated kernels.  On a GPU box the
drop-in is used directly (PYTHONPATH=dropin) and this file plays no part."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (_ROOT, os.path.join(_ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import conftest as _conftest  # noqa: E402
from curdleproofs_pie_b200 import runtime as _runtime  # noqa: E402

_runtime._install_library_for_tests(_runtime.CpgLib(_conftest.build_seam(), 0))
_spec = importlib.util.spec_from_file_location("_cpg_dropin_surface", os.path.join(_ROOT, "dropin", "py_arkworks_bls12381", "__init__.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
G1Point = _mod.G1Point
Scalar = _mod.Scalar

"""The reference's OWN test functions, unmodified, imported from /root/reference and run on top of the
drop-in `py_arkworks_bls12381` surface (CPU tier: host-emulated kernels via tests/seam_shim).  Skipped
where the reference is not mounted (the GPU box)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "curdleproofs")), reason="reference not mounted")

# the dir() listing test is pinned to CPython <= 3.10 and fails for any implementation on 3.12
SELECT = ("test_py_arkworks_bls12381_g1points or test_py_arkworks_bls12381_scalar or test_scalar_pow or test_utils_point_projective_to_bytes "
          "or test_same_scalar_arg or test_group_commit or test_shuffle_argument or test_tracker_opening_proof "
          "or test_whisk_interface_tracker_opening_proof or test_whisk_interface_shuffle_proof or test_serde")


def test_reference_tests_pass_on_the_dropin_surface(seam_lib):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "seam_shim"), os.path.join(REF, "curdleproofs"),
                                         os.path.join(REF, "merlin_transcripts"), ROOT])
    env["OMP_NUM_THREADS"] = "4"
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp",
                        os.path.join(REF, "curdleproofs", "curdleproofs", "test_curdleproofs.py"), "-k", SELECT],
                       env=env, cwd="/tmp", capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail

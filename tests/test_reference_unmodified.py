"""The UNMODIFIED reference (baseline/_ref, installed by tools/install_reference.py from /root/reference; it travels to
the GPU box with the snapshot) on top of BOTH drop-ins: dropin/py_arkworks_bls12381 (G1Point / Scalar on libcpg.so) and
dropin/merlin_transcripts (MerlinTranscript on libcpg.so's STROBE/Keccak).

  GPU tier: the reference's WHOLE test file curdleproofs/test_curdleproofs.py (cp/test_curdleproofs.py:132-775: every
            sub-argument at n = 128 with its negative cases, the N = 64 shuffle argument, bad shuffle arguments, serde,
            opening proofs, the Whisk API) minus the dir() listing test that is pinned to CPython <= 3.10,
            plus its merlin tests; and a replay of the n = 128 golden fixture: the reference's own CurdleProofsProof.new
            emits the fixture's proof bytes on the B200, .verify its five verdicts.
  CPU tier: the same replay at N = 8 on the host-emulated kernels.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "curdleproofs", "test_curdleproofs.py")),
                                reason="baseline/_ref not installed (python tools/install_reference.py)")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), REF, ROOT])
    return env


def _replay(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_on_dropin.py")] + args, env=dict(os.environ), cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_replays_golden_on_host_emulation(seam_lib):
    out = _replay(["--case", "shuffle_N8_seed1234.json", "--repeat", "1", "--test-seam"], 900)
    assert out["backend"] == "host-emulation-test-seam" and out["merlin_module"].startswith("dropin")


@pytest.mark.gpu
@pytest.mark.parametrize("defer", [[], ["--defer-decode"]], ids=["eager-decode", "deferred-decode"])
def test_reference_replays_golden_n128_on_gpu(gpu_lib, defer):
    out = _replay(["--case", "shuffle_N128_seed4096.json", "--repeat", "1"] + defer, 900)
    assert out["backend"] == "cuda-sm_100a" and out["gpu_launches"] > 1000 and out["merlin_module"].startswith("dropin")
    print("unmodified reference on the B200 drop-in, n = 128: new %.2f s, verify %.2f s" % (out["CurdleProofsProof_new_s"], out["CurdleProofsProof_verify_s"]))


@pytest.mark.gpu
@pytest.mark.parametrize("defer", ["", "1"], ids=["eager-decode", "deferred-decode"])
def test_reference_whole_test_file_on_gpu(gpu_lib, defer):
    """Both decoding modes of the drop-in: the reference's tests expect no ValueError at a decoding call, so the whole
    file must pass with deferral on as well (CPG_DROPIN_DEFER_DECODE=1)."""
    env = _env()
    env["CPG_DROPIN_DEFER_DECODE"] = defer
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp",
                        os.path.join(REF, "curdleproofs", "test_curdleproofs.py"), os.path.join(REF, "merlin_transcripts", "test_merlin.py"),
                        "-k", "not test_py_arkworks_bls12381_api",
                        "--durations", "0"],
                       env=env, cwd="/tmp", capture_output=True, text=True, timeout=2400)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "deselected" in r.stdout, tail
    print(tail)

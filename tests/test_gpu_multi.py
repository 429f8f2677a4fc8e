"""GPU tier, boxes with >= 2 GPUs (skipped on one): the in-library communicator end to end, one process per GPU and no
torch - tools/comm_check.py: cpg_g1_msm_sharded == the single-GPU MSM == the oracle; ONE proof split over the ranks
(cpg_prover_create_sharded / cpg_verifier_create_sharded) gives the reference's golden proof bytes and verdicts at
N = 16 / 128 and its digests at N = 1024; all-gather / max-reduce / verdict gather helpers."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def test_sharded_msm_and_sharded_proof_over_two_gpus(gpu_lib):
    if int(gpu_lib.c.cpg_device_count()) < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "comm_check.py"), "--gpus", "2", "--terms", "65536", "--port", "29631"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["gpus"] == 2 and out["nccl_version"] > 0
    print(out)

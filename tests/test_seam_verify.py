"""CPU tier for the batched verifier: host transcript + Fr algebra (native C++) with the group
arithmetic on the host-emulated kernels, against the reference's golden verdicts."""
import ctypes

import verify_cases as vc


import pytest


@pytest.mark.parametrize("on_device", [True, False])
def test_verify_batch_N8(seam_lib, on_device):
    vc.check_batch(seam_lib, "shuffle_N8_seed1234.json", transcript_on_device=on_device, fixed_window=4)


@pytest.mark.parametrize("on_device", [True, False])
def test_verify_batch_N16_two_copies(seam_lib, on_device):
    vc.check_batch(seam_lib, "shuffle_N16_seed77.json", copies=2, window=4, transcript_on_device=on_device, fixed_window=5)


@pytest.mark.parametrize("on_device,group", [(True, 2), (False, 4)])
def test_verify_batch_cross_proof_groups(seam_lib, on_device, group):
    """groups of proofs share one aggregated MSM, failing groups fall back to per-proof checks;
    the verdict of every lane (honest and corrupted variants side by side) still equals the reference's"""
    vc.check_batch(seam_lib, "shuffle_N8_seed1234.json", transcript_on_device=on_device, fixed_window=4, group=group)

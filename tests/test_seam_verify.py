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


@pytest.mark.parametrize("on_device,group", [(True, 1), (False, 1), (True, 2)])
def test_tracker_cache_changes_no_verdict(seam_lib, on_device, group):
    """decompressed trackers cached by their 48-byte encoding (hits, in-batch duplicates, malformed
    encodings, table overflow) - verdicts equal the uncached ones lane by lane"""
    vc.check_cache(seam_lib, "shuffle_N8_seed1234.json", transcript_on_device=on_device, group=group)


def test_adaptive_group_size_follows_the_failure_rate(seam_lib):
    """group = 0: the library re-picks the group size after every batch (argmin of agg[G] + 1 - (1 - p)^G):
    all-valid batches drive it to 64, a batch where every second proof is bad drives it down to 2"""
    import shuffle_cases as sc
    from curdleproofs_pie_b200 import whisk

    case = sc.load_case("shuffle_N8_seed1234.json")
    vs = vc.variants(case)
    honest, bad = vs[0], vs[1]
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), case["N"] - 4, fixed_window=4, lib=seam_lib, group=0)
    assert ver.group() == 16
    B = 32
    assert ver.verify([honest[1]] * B, [honest[2]] * B) == [True] * B
    assert ver.group() == 64 and ver.rechecked() == 0
    ins = [honest[1] if i % 2 else bad[1] for i in range(B)]
    prs = [honest[2] if i % 2 else bad[2] for i in range(B)]
    want = [bool(i % 2) for i in range(B)]
    assert ver.verify(ins, prs) == want            # 64 > B: this batch runs with groups of 32 - one group, which fails
    assert ver.rechecked() == B and ver.group() == 4       # one failing group of 32: p >= 1/32
    assert ver.verify(ins, prs) == want and ver.group() == 2 and ver.rechecked() == B   # 8 of 8 groups of 4 fail: p >= 1/4
    assert ver.verify(ins, prs) == want and ver.group() == 2
    assert ver.verify([honest[1]] * 5, [honest[2]] * 5) == [True] * 5 and ver.rechecked() == 0
    ver.close()


def test_verifier_refuses_to_exist_without_entropy(seam_lib, monkeypatch):
    """the batching weights are only sound while unpredictable: no OS randomness -> cpg_verifier_create fails
    (round 1 silently fell back to a constant secret)"""
    import shuffle_cases as sc
    from curdleproofs_pie_b200 import runtime as rt
    from curdleproofs_pie_b200 import whisk

    case = sc.load_case("shuffle_N8_seed1234.json")
    monkeypatch.setenv("CPG_TEST_NO_ENTROPY", "1")
    with pytest.raises(rt.CpgError, match="random"):
        whisk.BatchVerifier(bytes.fromhex(case["crs"]), case["N"] - 4, fixed_window=4, lib=seam_lib)
    monkeypatch.delenv("CPG_TEST_NO_ENTROPY")
    whisk.BatchVerifier(bytes.fromhex(case["crs"]), case["N"] - 4, fixed_window=4, lib=seam_lib).close()


@pytest.mark.parametrize("ranks,on_device", [(2, False), (3, True), (8, False)])
def test_one_proof_split_over_ranks(seam_lib, monkeypatch, ranks, on_device):
    """every proof's MSM terms split over `ranks` blocks (emulated in-process), 2 partial sums per proof and block added
    up: honest and corrupted variants still get the reference's verdicts"""
    monkeypatch.setenv("CPG_TEST_VIRTUAL_RANKS", str(ranks))
    vc.check_batch(seam_lib, "shuffle_N8_seed1234.json", transcript_on_device=on_device, fixed_window=4, sharded=True)

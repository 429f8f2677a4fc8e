"""CPU tier for the batched prover: per-round step kernels and MSMs on the host-emulated kernels,
proof bytes against the reference's golden fixtures."""
import prove_cases as pc


import pytest


@pytest.mark.parametrize("transcript", ["host", "device"])
def test_prove_bytes_N8(seam_lib, transcript):
    pc.check_prove(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4, transcript=transcript)


@pytest.mark.parametrize("transcript", ["host", "device"])
def test_prove_bytes_N16_two_lanes(seam_lib, transcript):
    pc.check_prove(seam_lib, "shuffle_N16_seed77.json", copies=2, fixed_window=5, window=3, transcript=transcript)


def test_prove_bytes_N64_chunked_reductions(seam_lib):
    """n = 64: the two-pass sums / prefix products run over 4 chunks of 16 (n <= 16 has a single chunk)"""
    pc.check_prove(seam_lib, "shuffle_N64_seed2024.json", fixed_window=4, transcript="device")


def test_prove_then_verify_N8(seam_lib):
    pc.check_prove_then_verify(seam_lib, "shuffle_N8_seed1234.json", B=3, fixed_window=4)


def test_prove_sub_batches_on_stream_lanes(seam_lib):
    """the batch split over 3 lanes of unequal size (5 = 1 + 2 + 2): bytes still equal the reference's"""
    pc.check_prove(seam_lib, "shuffle_N8_seed1234.json", copies=5, fixed_window=4, lanes=(3, 1))
    pc.check_prove_then_verify(seam_lib, "shuffle_N8_seed1234.json", B=4, fixed_window=4, lanes=(2, 2))


def test_prove_tracker_msms_bucket_method_and_small_tables(seam_lib):
    """the T / U MSMs either through per-base tables of multiples (default, window 6) or through the bucket
    method (0); a 3-bit table exercises the top-digit edge of the signed recoding"""
    pc.check_prove(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4, table_window=0)
    pc.check_prove(seam_lib, "shuffle_N16_seed77.json", fixed_window=5, table_window=3)


def test_prove_rejects_non_canonical_k(seam_lib):
    pc.check_rejects_non_canonical_k(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4)


def test_prove_with_randomness_continued_in_c(seam_lib):
    pc.check_prove_drawn(seam_lib, "shuffle_N8_seed1234.json", B=2, fixed_window=4)


def test_prove_with_identity_trackers_matches_the_oracle(seam_lib):
    pc.check_prove_with_identity_trackers(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4)
    pc.check_prove_with_identity_trackers(seam_lib, "shuffle_N16_seed77.json", fixed_window=5, table_window=0)


def test_prove_rejects_bad_perm_and_blinders(seam_lib):
    pc.check_rejects_bad_perm_and_blinders(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4)


def test_whisk_api_generate_then_validate(seam_lib):
    pc.check_whisk_api_roundtrip(seam_lib, "shuffle_N8_seed1234.json", B=2)


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_one_proof_split_over_ranks(seam_lib, monkeypatch, ranks):
    """BASELINE config 5's decomposition: the leaves of ONE proof split over `ranks` blocks (CRS bases through a table
    per block, trackers through the bucket method over sub-ranges of the coefficient rows), partial sums added up -
    the communicator is emulated in-process (CPG_TEST_VIRTUAL_RANKS), everything else is the product's own code.
    Bytes still equal the reference's; ranks = 8 > ell = 4 leaves some blocks without trackers."""
    monkeypatch.setenv("CPG_TEST_VIRTUAL_RANKS", str(ranks))
    pc.check_prove(seam_lib, "shuffle_N8_seed1234.json", fixed_window=4, sharded=True, transcript="host")
    if ranks == 3:
        pc.check_prove(seam_lib, "shuffle_N16_seed77.json", copies=2, fixed_window=5, window=3, sharded=True, transcript="device")

"""GPU tier: parity tests proper.  Every case calls the CUDA kernels through the C ABI
(include/cpg.h) and compares compressed bytes with the CPU oracle on the same seeded inputs;
full-size shapes are covered by size-independent properties (linearity, splitting, round trips).
Run on a B200: python -m pytest tests -m gpu"""
import os
import importlib
import random

import pytest

import dropin_cases as dc
import parity_cases as pc
import shuffle_cases as sc
from curdleproofs_pie_b200 import runtime as rt

pytestmark = pytest.mark.gpu


def test_roundtrip(gpu_lib, cref):
    pc.case_roundtrip(gpu_lib, cref, 300)


def test_group_law(gpu_lib, cref):
    pc.case_group_law(gpu_lib, cref, 200)


def test_fold(gpu_lib, cref):
    pc.case_fold(gpu_lib, cref, 8, 64)


@pytest.mark.parametrize("B,n,window,shared", [
    (64, 128, 0, False), (32, 128, 0, True), (8, 627, 0, False), (3, 1, 0, False), (5, 7, 2, False),
    (4, 124, 4, False), (4, 64, 6, False), (2, 1000, 9, False), (1, 4096, 0, False), (16, 33, 7, True),
])
def test_msm(gpu_lib, cref, B, n, window, shared):
    pc.case_msm(gpu_lib, cref, B, n, window, shared)


@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("B,n,window,shared", [(64, 128, 0, False), (8, 627, 7, False), (5, 7, 2, False), (16, 33, 7, True)])
def test_msm_both_pipelines(gpu_lib, cref, B, n, window, shared, path):
    gpu_lib.check(gpu_lib.c.cpg_msm_force_path(path))
    try:
        pc.case_msm(gpu_lib, cref, B, n, window, shared)
        pc.case_msm(gpu_lib, cref, 6, 40, 4, shared=False, edge=True)
    finally:
        gpu_lib.check(gpu_lib.c.cpg_msm_force_path(0))


def test_msm_edges(gpu_lib, cref):
    pc.case_msm(gpu_lib, cref, 6, 40, 4, shared=False, edge=True)
    pc.case_msm(gpu_lib, cref, 5, 128, 0, shared=True, edge=True)


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("path,n,window", [(2, 5000, 0), (1, 600, 7), (2, 20000, 12), (2, 64, 0)])
def test_msm_skewed_digits(gpu_lib, cref, path, n, window, accumulate):
    gpu_lib.check(gpu_lib.c.cpg_msm_force_path(path))
    gpu_lib.check(gpu_lib.c.cpg_msm_set_accumulate(accumulate))
    try:
        pc.case_msm_skewed(gpu_lib, cref, n, window)
    finally:
        gpu_lib.check(gpu_lib.c.cpg_msm_force_path(0))
        gpu_lib.check(gpu_lib.c.cpg_msm_set_accumulate(0))


@pytest.mark.parametrize("B,n,window,shared", [(64, 128, 0, False), (8, 627, 7, False), (2, 3000, 9, False), (1, 40000, 0, False)])
def test_msm_batched_affine_accumulation_matches(gpu_lib, cref, B, n, window, shared):
    """cpg_msm_set_accumulate(1): bucket sums by batched affine additions (pairwise trees, 32 additions per shared
    inversion, warp-converged) - an alternative to the default mixed-XYZZ accumulation, same bytes"""
    gpu_lib.check(gpu_lib.c.cpg_msm_set_accumulate(1))
    try:
        if n > 5000:
            pc.case_msm_large(gpu_lib, cref, n, window)
        else:
            pc.case_msm(gpu_lib, cref, B, n, window, shared)
            pc.case_msm(gpu_lib, cref, 6, 40, 4, shared=False, edge=True)
    finally:
        gpu_lib.check(gpu_lib.c.cpg_msm_set_accumulate(0))


def test_msm_empty(gpu_lib):
    out = gpu_lib.msm_batched(gpu_lib.alloc(96), 0, gpu_lib.alloc(32), 2, 0)
    assert gpu_lib.is_identity(out, 2) == [1, 1]


def test_fixed_base(gpu_lib, cref):
    pc.case_fixed(gpu_lib, cref, 8, 131, 8)
    pc.case_fixed(gpu_lib, cref, 3, 5, 4)
    pc.case_fixed(gpu_lib, cref, 2, 1300, 5)      # few MSMs over many bases: the bases are split over threads (one large proof)
    pc.case_fixed(gpu_lib, cref, 4, 9, 12, with_identity=True)     # 2048-entry rows: 32 segments per row, identity rows skipped
    pc.case_fixed(gpu_lib, cref, 2, 3, 16)                         # the bench's table window


def test_fr(gpu_lib):
    pc.case_fr(gpu_lib, 1000)


def test_msm_full_size_properties(gpu_lib, cref):
    """B = 4096 MSMs of n = 128 (BASELINE config 3's MSM shape): linearity in the scalars and
    splitting over the bases, checked on the device for every lane; a sample of lanes is also
    compared with the oracle."""
    lib = gpu_lib
    B, n = 4096, 128
    rng = random.Random(99)
    ks = lib.upload(rt.scalars_to_bytes(rng.randrange(rt.R_ORDER) for _ in range(n)))
    gen = lib.generator()
    gens = lib.alloc(n * rt.JAC)
    for i in range(n):
        lib.check(lib.c.cpg_d2d(gens.ptr + i * rt.JAC, gen.ptr, rt.JAC))
    base_aff = lib.jac_to_aff(lib.mul(gens, ks, n), n)              # shared base vector P_i = k_i G
    a = [rng.randrange(rt.R_ORDER) for _ in range(B * n)]
    b = [rng.randrange(rt.R_ORDER) for _ in range(B * n)]
    da, db = lib.upload(rt.scalars_to_bytes(a)), lib.upload(rt.scalars_to_bytes(b))
    dab = lib.fr_op("add", da, db, B * n)
    ma = lib.msm_batched(base_aff, 0, da, B, n)
    mb = lib.msm_batched(base_aff, 0, db, B, n)
    mab = lib.msm_batched(base_aff, 0, dab, B, n, window=7)         # different window on purpose
    assert lib.eq(lib.add(ma, mb, B), mab, B) == [1] * B
    # fixed-base path agrees with the bucket path
    table = lib.fixed_table(base_aff, n, 8)
    assert lib.eq(lib.msm_fixed_batched(table, da, B), ma, B) == [1] * B
    # oracle on a few lanes
    enc = pc.split48(lib.compress_aff(base_aff, n))
    blobs = [cref.decompress(e, False) for e in enc]
    got = pc.split48(lib.compress_jac(ma, B))
    for lane in (0, 1, 2047, 4095):
        assert got[lane] == cref.compress(cref.msm(blobs, a[lane * n:(lane + 1) * n]))


@pytest.fixture(scope="module")
def dropin(gpu_lib):
    return importlib.import_module("py_arkworks_bls12381")


def test_msm_pipelined_slices_bit_exact(gpu_lib):
    """The opt-in pipelined single MSM (CPG_MSM_SLICES window slices on a low- and a high-priority stream, Horner pass in
    segments that continue from the partial result): same bytes as the oracle (tools/msm_trace.py asserts it)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for slices in ("2", "4"):
        env = dict(os.environ, CPG_MSM_SLICES=slices, CPG_MSM_PIPE_MIN_N="1")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "msm_trace.py"), "17", "/tmp/msm_trace_test.txt"], env=env, cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "timeline in" in r.stdout, (r.stdout + r.stderr)[-1500:]
        with open("/tmp/msm_trace_test.txt") as f:
            names = [ln.split()[0] for ln in f if ln.strip()]
        assert names.count("BucketAccumulate") == int(slices) and names.count("HornerJacCoop") == int(slices), names


def test_empty_batches(gpu_lib):
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shuffle_N8_seed1234.json")) as f:
        case = json.load(f)
    pc.case_empty(gpu_lib, case["crs"], case["N"] - 4)


def test_dropin_surface(dropin):
    dc.surface_kats(dropin)


def test_dropin_deferred_decoding(dropin, cref):
    dc.deferred_decoding(dropin, cref, n=300)


def test_dropin_multiexp(dropin, cref):
    dc.multiexp_matches_oracle(dropin, cref, 128)
    dc.multiexp_matches_oracle(dropin, cref, 1024, seed=12)


@pytest.mark.parametrize("name", ["shuffle_N8_seed1234.json", "shuffle_N16_seed77.json", "shuffle_N64_seed2024.json", "shuffle_N128_seed4096.json"])
def test_dropin_prove_bytes_equal_reference(dropin, name):
    sc.check_prove_matches_golden(dropin, sc.load_case(name))


@pytest.mark.parametrize("name", ["shuffle_N8_seed1234.json", "shuffle_N64_seed2024.json"])
def test_dropin_verify_verdicts_equal_reference(dropin, name):
    sc.check_verify_matches_golden(dropin, sc.load_case(name))


# ---- batched verifier (cpg_verify_batch) against the reference's golden verdicts ----
import verify_cases as vc  # noqa: E402


@pytest.mark.parametrize("name,copies,window", [("shuffle_N8_seed1234.json", 1, 0), ("shuffle_N64_seed2024.json", 3, 0),
                                                 ("shuffle_N128_seed4096.json", 40, 0), ("shuffle_N128_seed4096.json", 2, 5)])
def test_verify_batch(gpu_lib, name, copies, window):
    vc.check_batch(gpu_lib, name, copies=copies, window=window, transcript_on_device=True)
    vc.check_batch(gpu_lib, name, copies=min(copies, 2), window=window, transcript_on_device=False)


@pytest.mark.parametrize("name,copies,group", [("shuffle_N8_seed1234.json", 1, 2), ("shuffle_N16_seed77.json", 12, 8),
                                               ("shuffle_N128_seed4096.json", 6, 16), ("shuffle_N64_seed2024.json", 40, 32)])
def test_verify_batch_cross_proof_groups(gpu_lib, name, copies, group):
    vc.check_batch(gpu_lib, name, copies=copies, transcript_on_device=True, group=group)
    if copies <= 6:
        vc.check_batch(gpu_lib, name, copies=2, transcript_on_device=False, group=group)


@pytest.mark.parametrize("name,lg,group,streams,copies", [("shuffle_N8_seed1234.json", 12, 1, 0, 3), ("shuffle_N128_seed4096.json", 16, 1, 4, 36),
                                                          ("shuffle_N128_seed4096.json", 14, 8, 4, 36), ("shuffle_N64_seed2024.json", 16, 1, 1, 3)])
def test_tracker_cache_changes_no_verdict(gpu_lib, name, lg, group, streams, copies):
    vc.check_cache(gpu_lib, name, log2_slots=lg, group=group, streams=streams, copies=copies)
    vc.check_cache(gpu_lib, name, log2_slots=lg, transcript_on_device=False, group=group)


@pytest.mark.parametrize("on_device", [1, 2])
def test_transcript_kats_on_device(gpu_lib, on_device):
    """The product's STROBE / Merlin / Keccak in a real kernel - on one thread (1) and on one warp in lock-step with the
    warp-cooperative permutation (2) - against the reference's known answers (mt/test_merlin.py:18,29,40) and the oracle"""
    import test_host_transcript as tht

    tht.strobe_conformance(gpu_lib, on_device)
    tht.merlin_simple(gpu_lib, on_device)
    for seed in range(6):
        tht.merlin_random_scripts(gpu_lib, on_device, seed, rounds=20)


@pytest.mark.parametrize("name,copies,group", [("shuffle_N8_seed1234.json", 1, 1), ("shuffle_N16_seed77.json", 3, 2), ("shuffle_N64_seed2024.json", 2, 1),
                                               ("shuffle_N128_seed4096.json", 20, 1), ("shuffle_N128_seed4096.json", 40, 8)])
def test_verify_batch_warp_per_proof_transcript(gpu_lib, name, copies, group):
    """cpg_verifier_set_transcript(3): one WARP per proof (what mode 2 picks for tens to a few thousand proofs)"""
    vc.check_batch(gpu_lib, name, copies=copies, transcript_on_device="warp", group=group)
    vc.check_cache(gpu_lib, name, log2_slots=14, transcript_on_device="warp", group=group)


def test_verify_replay_matches(gpu_lib):
    import ctypes

    from curdleproofs_pie_b200 import whisk

    case = sc.load_case("shuffle_N128_seed4096.json")
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), 124, lib=gpu_lib)
    vs = vc.variants(case) * 8
    got = ver.verify([v[1] for v in vs], [v[2] for v in vs])
    out = ctypes.create_string_buffer(len(vs))
    gpu_lib.check(gpu_lib.c.cpg_verify_replay_device(ver.handle, out))
    # every lane has the right length, so no verdict was decided on the host: the replay of the device side must
    # reproduce the whole bitmap, rejecting lanes included (this is the function bench.py's `value` times)
    assert [bool(x) for x in out.raw[:len(vs)]] == got
    assert got == [v[3] for v in vs]
    for group in (1, 8):
        ver.set_group(group)
        assert ver.verify([v[1] for v in vs], [v[2] for v in vs]) == got
        gpu_lib.check(gpu_lib.c.cpg_verify_replay_device(ver.handle, out))
        assert [bool(x) for x in out.raw[:len(vs)]] == got
    ver.close()


# ---- batched prover (cpg_prove_batch): proof bytes equal the reference's under the fixture's seed ----
import prove_cases as prc  # noqa: E402


@pytest.mark.parametrize("name,copies", [("shuffle_N8_seed1234.json", 1), ("shuffle_N16_seed77.json", 2), ("shuffle_N64_seed2024.json", 2),
                                          ("shuffle_N128_seed4096.json", 3)])
def test_prove_batch_bytes_equal_reference(gpu_lib, name, copies):
    prc.check_prove(gpu_lib, name, copies=copies)


def test_prove_then_verify_full_size(gpu_lib):
    prc.check_prove_then_verify(gpu_lib, "shuffle_N128_seed4096.json", B=96)


@pytest.mark.parametrize("table_window", [0, 4, 7])
def test_prove_tracker_msm_table_windows(gpu_lib, table_window):
    prc.check_prove(gpu_lib, "shuffle_N128_seed4096.json", copies=3, table_window=table_window)


def test_prove_with_randomness_continued_in_c(gpu_lib):
    prc.check_prove_drawn(gpu_lib, "shuffle_N128_seed4096.json", B=5)


def test_prove_with_identity_trackers_matches_the_oracle(gpu_lib):
    prc.check_prove_with_identity_trackers(gpu_lib, "shuffle_N16_seed77.json")
    prc.check_prove_with_identity_trackers(gpu_lib, "shuffle_N64_seed2024.json", table_window=0)


def test_prove_rejects_non_canonical_k(gpu_lib):
    prc.check_rejects_non_canonical_k(gpu_lib, "shuffle_N16_seed77.json")


def test_prove_rejects_bad_perm_and_blinders(gpu_lib):
    prc.check_rejects_bad_perm_and_blinders(gpu_lib, "shuffle_N16_seed77.json")


def test_whisk_api_generate_then_validate(gpu_lib):
    prc.check_whisk_api_roundtrip(gpu_lib, "shuffle_N64_seed2024.json", B=5)


def test_prove_sub_batches_on_stream_lanes(gpu_lib):
    prc.check_prove(gpu_lib, "shuffle_N64_seed2024.json", copies=7, lanes=(3, 2))
    prc.check_prove_then_verify(gpu_lib, "shuffle_N128_seed4096.json", B=96, lanes=(4, 16))
    prc.check_prove_then_verify(gpu_lib, "shuffle_N8_seed1234.json", B=600)          # default: 2 lanes from 512 proofs


@pytest.mark.parametrize("n,window", [(5000, 0), (3000, 10), (1 << 16, 13), (100000, 0)])
def test_msm_large_and_window_slices(gpu_lib, cref, n, window):
    pc.case_msm_large(gpu_lib, cref, n, window)


# ---- large shuffles (BASELINE config 5 and a test-sized stand-in): digests of the reference's outputs ----
import large_cases as lc  # noqa: E402


def test_large_shuffle_N1024_bytes_and_verdicts_equal_reference(gpu_lib):
    lc.check_large(gpu_lib, "large_N1024_seed6024.json", fixed_window=8)


def test_large_shuffle_N16384_config5_single_gpu(gpu_lib):
    """BASELINE config 5 on one GPU: 7808-byte proof, digests of the unmodified reference's outputs"""
    lc.check_large(gpu_lib, "large_N16384_seed21384.json", fixed_window=8)

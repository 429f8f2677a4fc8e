"""CPU tier: the kernels' per-thread code (host-emulated, see tests/conftest.py::build_seam) driven
through the same C ABI and launch logic as on the GPU, checked bit-for-bit against the oracle.
Sizes are small; the GPU tier (tests/test_gpu_parity.py) repeats the cases at full size."""
import os
import pytest

import parity_cases as pc


def test_roundtrip(seam_lib, cref):
    pc.case_roundtrip(seam_lib, cref, 12)


def test_group_law(seam_lib, cref):
    pc.case_group_law(seam_lib, cref, 12)


def test_fold(seam_lib, cref):
    pc.case_fold(seam_lib, cref, 2, 4)


@pytest.mark.parametrize("B,n,window,shared", [(3, 16, 0, False), (2, 33, 5, True), (1, 1, 0, False), (2, 7, 2, False), (1, 40, 8, False)])
def test_msm(seam_lib, cref, B, n, window, shared):
    pc.case_msm(seam_lib, cref, B, n, window, shared)


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("B,n,window,shared", [(3, 16, 0, False), (2, 33, 5, True), (2, 7, 2, False), (1, 40, 8, False), (2, 200, 3, False)])
def test_msm_both_pipelines(seam_lib, cref, B, n, window, shared, path, accumulate):
    """path 1 = one thread per (msm, window) (the batched-proofs pipeline), 2 = per-term threads with the
    level-wise reduction (single / large MSMs); by shape these small batches would all take path 2.
    accumulate 0 = mixed XYZZ additions, 1 = batched affine additions (pairwise trees, shared inversions)"""
    seam_lib.check(seam_lib.c.cpg_msm_force_path(path))
    seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(accumulate))
    try:
        pc.case_msm(seam_lib, cref, B, n, window, shared)
        # zero scalars, k = r - 1, identity bases, P and -P / the same base twice in one bucket
        pc.case_msm(seam_lib, cref, 4, 12, 4, shared=False, edge=True)
        pc.case_msm(seam_lib, cref, 3, 12, 3, shared=True, edge=True)
        pc.case_msm(seam_lib, cref, 2, 90, 2, shared=True, edge=True)        # 2 buckets per window: long lists of equal / opposite points
    finally:
        seam_lib.check(seam_lib.c.cpg_msm_force_path(0))
        seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(0))


def test_msm_edges(seam_lib, cref):
    pc.case_msm(seam_lib, cref, 4, 12, 4, shared=False, edge=True)
    pc.case_msm(seam_lib, cref, 3, 12, 3, shared=True, edge=True)


def test_empty_batches(seam_lib):
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shuffle_N8_seed1234.json")) as f:
        case = json.load(f)
    pc.case_empty(seam_lib, case["crs"], case["N"] - 4)


def test_msm_empty(seam_lib, cref):
    out = seam_lib.msm_batched(seam_lib.alloc(96), 0, seam_lib.alloc(32), 2, 0)
    assert seam_lib.is_identity(out, 2) == [1, 1]


def test_fixed_base(seam_lib, cref):
    pc.case_fixed(seam_lib, cref, 2, 5, 4)
    # rows longer than one 64-entry segment (NB = 128 / 256: 2 / 4 segments, each with its own shared inversion),
    # and an identity base whose rows are all identities
    pc.case_fixed(seam_lib, cref, 2, 3, 8, with_identity=True)
    pc.case_fixed(seam_lib, cref, 1, 2, 9)


def test_fr(seam_lib):
    pc.case_fr(seam_lib, 16)


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("n,window", [(2100, 5), (300, 10), (2500, 9), (600, 16), (3000, 0)])
def test_msm_large_and_window_slices(seam_lib, cref, n, window, accumulate):
    seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(accumulate))
    try:
        pc.case_msm_large(seam_lib, cref, n, window)
    finally:
        seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(0))


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("path,n,window", [(2, 700, 6), (1, 300, 4), (2, 64, 0)])
def test_msm_skewed_digits(seam_lib, cref, path, n, window, accumulate):
    seam_lib.check(seam_lib.c.cpg_msm_force_path(path))
    seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(accumulate))
    try:
        pc.case_msm_skewed(seam_lib, cref, n, window)
    finally:
        seam_lib.check(seam_lib.c.cpg_msm_force_path(0))
        seam_lib.check(seam_lib.c.cpg_msm_set_accumulate(0))

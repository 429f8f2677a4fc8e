"""Drive oracle/shuffle_ref.py (the per-proof restatement of the reference's shuffle argument) on an
arbitrary G1Point/Scalar surface - the oracle's or the CUDA drop-in's - and compare with the golden
fixtures generated from the UNMODIFIED reference (oracle/gen_golden.py)."""
import json
import os
import random

from oracle.shuffle_ref import N_BLINDERS, ShuffleRef

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def prove_like_reference(surface, case):
    """Replays the reference test's construction order under the case's seed; returns
    (ctx, crs, inputs..., proof_bytes)."""
    G1Point, Scalar = surface.G1Point, surface.Scalar
    ctx = ShuffleRef(G1Point, Scalar)
    random.seed(case["seed"])
    N = case["N"]
    ell = N - N_BLINDERS
    crs = ctx.make_crs(ell)
    perm = list(range(ell))
    random.shuffle(perm)
    k = ctx.rand()
    vec_R = [ctx.gen * ctx.rand() for _ in range(ell)]
    vec_S = [ctx.gen * ctx.rand() for _ in range(ell)]
    vec_T, vec_U, M, m_bl = ctx.shuffle_and_commit(crs, vec_R, vec_S, perm, k)
    proof = ctx.prove(crs, vec_R, vec_S, vec_T, vec_U, M, perm, k, m_bl)
    return ctx, crs, vec_R, vec_S, vec_T, vec_U, M, proof


def check_prove_matches_golden(surface, case):
    ctx, crs, vec_R, vec_S, vec_T, vec_U, M, proof = prove_like_reference(surface, case)
    assert perm_ok(case)
    assert ctx.crs_to_bytes(crs).hex() == case["crs"]
    assert [ctx.pb(p).hex() for p in vec_T] == case["vec_T"]
    assert [ctx.pb(p).hex() for p in vec_U] == case["vec_U"]
    assert ctx.pb(M).hex() == case["M"]
    assert proof.hex() == case["proof"], "proof bytes differ from the reference's"


def perm_ok(case):
    return sorted(case["perm"]) == list(range(case["N"] - N_BLINDERS))


def check_verify_matches_golden(surface, case):
    """Verify the REFERENCE's proof bytes from the fixture (inputs decoded from wire bytes)."""
    G1Point, Scalar = surface.G1Point, surface.Scalar
    ctx = ShuffleRef(G1Point, Scalar)
    N = case["N"]
    ell = N - N_BLINDERS
    crs = ctx.crs_from_bytes(bytes.fromhex(case["crs"]), ell)
    dec = lambda lst: [G1Point.from_compressed_bytes_unchecked(bytes.fromhex(h)) for h in lst]  # noqa: E731
    R_, S_, T_, U_ = dec(case["vec_R"]), dec(case["vec_S"]), dec(case["vec_T"]), dec(case["vec_U"])
    M = G1Point.from_compressed_bytes_unchecked(bytes.fromhex(case["M"]))
    proof = bytes.fromhex(case["proof"])
    got = {
        "honest": ctx.is_valid(crs, R_, S_, T_, U_, M, proof),
        "swap_R_S": ctx.is_valid(crs, S_, R_, T_, U_, M, proof),
        "swap_T_U": ctx.is_valid(crs, R_, S_, U_, T_, M, proof),
        "wrong_M": ctx.is_valid(crs, R_, S_, T_, U_, M + M, proof),
        "rotated_T": ctx.is_valid(crs, R_, S_, T_[1:] + T_[:1], U_, M, proof),
    }
    assert got == case["verdicts"]

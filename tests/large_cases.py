"""Large shuffles (BASELINE config 5 and its test-sized stand-in) through the batched prover and verifier.
The fixtures (oracle/gen_golden_large.py) hold only the seed and SHA-256 digests of what the UNMODIFIED
reference produced; every input point is s_i * G with s_i drawn from Python's `random` in the reference's
order, so the inputs are rebuilt here on the device and pinned by their digests first."""
import hashlib
import json
import os
import random
import time

from curdleproofs_pie_b200 import runtime as rt
from curdleproofs_pie_b200 import whisk


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _scalar_mul_generator(lib, scalars):
    """compressed bytes of s_i * G, and the Jacobian points"""
    k = len(scalars)
    gen = lib.generator()
    gens = lib.alloc(k * rt.JAC)
    lib.check(lib.c.cpg_d2d(gens.ptr, gen.ptr, rt.JAC))
    have = 1
    while have < k:
        cnt = min(have, k - have)
        lib.check(lib.c.cpg_d2d(gens.ptr + have * rt.JAC, gens.ptr, cnt * rt.JAC))
        have += cnt
    pts = lib.mul(gens, lib.upload(rt.scalars_to_bytes(scalars)), k)
    return lib.compress_jac(pts, k), pts


def rebuild_inputs(lib, case, n_rand):
    """Replays cp/test_curdleproofs.py:576-593's construction order under the fixture's seed: CRS (ell + 4 + 3
    random points, then G_sum / H_sum), permutation, k, vec_R, vec_S, then the prover's blinders."""
    N = case["N"]
    ell = N - 4
    random.seed(case["seed"])
    draw = lambda: random.randint(1, rt.R_ORDER - 1)  # noqa: E731
    crs_sc = [draw() for _ in range(ell + 4 + 3)]
    perm = list(range(ell))
    random.shuffle(perm)
    k = draw()
    r_sc = [draw() for _ in range(ell)]
    s_sc = [draw() for _ in range(ell)]
    rand = b"".join(draw().to_bytes(32, "little") for _ in range(n_rand))
    # G_sum = sum vec_G, H_sum = sum vec_H  (cp/crs.py:19-36): s * G summed = (sum s) * G
    g_sum = sum(crs_sc[:ell]) % rt.R_ORDER
    h_sum = sum(crs_sc[ell:ell + 4]) % rt.R_ORDER
    crs48, _ = _scalar_mul_generator(lib, crs_sc + [g_sum, h_sum])
    pre48, _ = _scalar_mul_generator(lib, r_sc + s_sc)
    assert k == case["k"]
    assert sha(crs48) == case["crs_sha256"], "CRS bytes differ from the reference's"
    assert sha(pre48) == case["pre_sha256"], "vec_R | vec_S bytes differ from the reference's"
    return crs48, pre48, perm, k, rand


def check_large(lib, name, fixed_window=8, sharded=False):
    """prove: trackers and proof bytes equal the reference's; verify: the reference's verdicts on the honest
    and the swapped inputs.  Returns the timings."""
    case = load_case(name)
    N = case["N"]
    ell = N - 4
    t0 = time.perf_counter()
    prover = whisk.BatchProver(_crs_probe(lib, case), ell, fixed_window=fixed_window, lib=lib, sharded=sharded)
    crs48, pre48, perm, k, rand = rebuild_inputs(lib, case, prover.n_rand)
    t1 = time.perf_counter()
    (tu, proof), = prover.prove([pre48], [perm], [k], [rand])
    t2 = time.perf_counter()
    prover.close()
    assert sha(tu) == case["post_sha256"], "post-shuffle trackers differ from the reference's"
    assert proof[:48].hex() == case["M"]
    assert len(proof) - 48 == case["proof_len"]
    assert sha(proof[48:]) == case["proof_sha256"], "proof bytes differ from the reference's"
    ver = whisk.BatchVerifier(crs48, ell, fixed_window=fixed_window, lib=lib, sharded=sharded)
    w = 48 * ell
    R_, S_, T_, U_ = pre48[:w], pre48[w:], tu[:w], tu[w:]
    t3 = time.perf_counter()
    got = ver.verify([pre48 + tu, S_ + R_ + tu, pre48 + U_ + T_], [proof] * 3)
    t4 = time.perf_counter()
    ver.close()
    v = case["verdicts"]
    assert got == [v["honest"], v["swap_R_S"], v["swap_T_U"]], got
    return {"N": N, "setup_s": t1 - t0, "prove_s": t2 - t1, "verify3_s": t4 - t3}


def _crs_probe(lib, case):
    """CRS wire bytes from the seed alone (needed before the prover exists to size its tables)."""
    N = case["N"]
    ell = N - 4
    random.seed(case["seed"])
    sc_ = [random.randint(1, rt.R_ORDER - 1) for _ in range(ell + 4 + 3)]
    g_sum = sum(sc_[:ell]) % rt.R_ORDER
    h_sum = sum(sc_[ell:ell + 4]) % rt.R_ORDER
    crs48, _ = _scalar_mul_generator(lib, sc_ + [g_sum, h_sum])
    return crs48

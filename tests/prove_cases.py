"""Batched prover (cpg_prove_batch) against the golden fixtures produced by the UNMODIFIED reference:
with the blinders drawn from Python's `random` in the reference's order under the fixture's seed, the
post-shuffle trackers, M and the proof wire bytes must equal the reference's bit for bit."""
import random

import shuffle_cases as sc
from curdleproofs_pie_b200 import runtime as rt
from curdleproofs_pie_b200 import whisk


def replay_rng(case, prover):
    """Re-create the reference test's RNG stream (oracle/gen_golden.py::one_case) up to the prover's draws."""
    N = case["N"]
    ell = N - 4
    random.seed(case["seed"])
    for _ in range(ell + 4 + 3):                    # CurdleproofsCrs.new: ell + n_bl + 3 random points
        random.randint(1, rt.R_ORDER - 1)
    perm = list(range(ell))
    random.shuffle(perm)
    k = random.randint(1, rt.R_ORDER - 1)
    for _ in range(2 * ell):                        # vec_R, vec_S
        random.randint(1, rt.R_ORDER - 1)
    rand = prover.draw_randomness(random)           # m_bl, then CurdleProofsProof.new's 3n + 9 draws
    assert perm == case["perm"] and k == case["k"]
    return perm, k, rand


def check_prove(lib, name, copies=1, fixed_window=0, window=0, lanes=None, table_window=None, transcript=None, sharded=False):
    """transcript: None (library default: by batch size), "host" or "device" - where the Fiat-Shamir step of every
    round runs; the Fr vector kernels are the same either way"""
    case = sc.load_case(name)
    ell = case["N"] - 4
    prover = whisk.BatchProver(bytes.fromhex(case["crs"]), ell, fixed_window=fixed_window, lib=lib, sharded=sharded)
    if transcript is not None:
        prover.set_transcript(transcript)
    if lanes:
        prover.set_lanes(*lanes)
    if table_window is not None:
        prover.set_table_window(table_window)
    if window:
        prover.set_window(window)
    perm, k, rand = replay_rng(case, prover)
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    res = prover.prove([pre] * copies, [perm] * copies, [k] * copies, [rand] * copies)
    want_tu = b"".join(bytes.fromhex(h) for h in case["vec_T"] + case["vec_U"])
    want_proof = bytes.fromhex(case["M"]) + bytes.fromhex(case["proof"])
    for tu, proof in res:
        assert tu == want_tu, "post-shuffle trackers differ from the reference's"
        assert proof[:48] == want_proof[:48], "M differs"
        assert proof == want_proof, "proof bytes differ from the reference's (first diff at %d)" % next(i for i in range(len(proof)) if proof[i] != want_proof[i])
    prover.close()
    return res


def check_prove_then_verify(lib, name, B=3, fixed_window=0, lanes=None):
    """Fresh randomness per lane (different permutations and k): every proof must verify, and must stop
    verifying when paired with another lane's trackers."""
    case = sc.load_case(name)
    ell = case["N"] - 4
    crs = bytes.fromhex(case["crs"])
    prover = whisk.BatchProver(crs, ell, fixed_window=fixed_window, lib=lib)
    if lanes:
        prover.set_lanes(*lanes)
    rng = random.Random(2718)
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    perms, ks, rands = [], [], []
    for _ in range(B):
        p = list(range(ell)); rng.shuffle(p)
        perms.append(p); ks.append(rng.randint(1, rt.R_ORDER - 1)); rands.append(prover.draw_randomness(rng))
    res = prover.prove([pre] * B, perms, ks, rands)
    prover.close()
    ver = whisk.BatchVerifier(crs, ell, fixed_window=fixed_window, lib=lib)
    inputs = [pre + tu for tu, _ in res]
    proofs = [pr for _, pr in res]
    assert ver.verify(inputs, proofs) == [True] * B
    assert ver.verify(inputs, proofs[1:] + proofs[:1]) == [False] * B
    ver.close()


def check_rejects_non_canonical_k(lib, name, fixed_window=0):
    """k >= r is not a Scalar (Scalar.from_le_bytes raises in the reference): that lane is flagged, the others prove"""
    import pytest

    case = sc.load_case(name)
    ell = case["N"] - 4
    prover = whisk.BatchProver(bytes.fromhex(case["crs"]), ell, fixed_window=fixed_window, lib=lib)
    perm, k, rand = replay_rng(case, prover)
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    ks = k.to_bytes(32, "little") + rt.R_ORDER.to_bytes(32, "little") + (2**256 - 1).to_bytes(32, "little")
    tu, pr, st = prover.prove_raw(pre * 3, perm * 3, ks, rand * 3, 3)
    assert list(st) == [0, 1, 1]
    assert pr[:prover.proof_len] == bytes.fromhex(case["M"]) + bytes.fromhex(case["proof"])
    with pytest.raises(ValueError):
        prover.prove([pre], [perm], [rt.R_ORDER], [rand])
    prover.close()


def check_prove_drawn(lib, name, B=3, fixed_window=0):
    """prove_drawn (randomness continued in C from the caller's `random` state) == prove() on values drawn by
    Python from a twin generator; the proofs verify"""
    case = sc.load_case(name)
    ell = case["N"] - 4
    crs = bytes.fromhex(case["crs"])
    prover = whisk.BatchProver(crs, ell, fixed_window=fixed_window, lib=lib)
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    a, b = random.Random(31337), random.Random(31337)
    got = prover.prove_drawn([pre] * B, a)
    perms, ks, rands = [], [], []
    for _ in range(B):
        p = list(range(ell)); b.shuffle(p)
        perms.append(p); ks.append(b.randint(1, rt.R_ORDER - 1)); rands.append(prover.draw_randomness(b))
    want = prover.prove([pre] * B, perms, ks, rands)
    assert got == want
    assert a.random() == b.random()
    prover.close()
    ver = whisk.BatchVerifier(crs, ell, fixed_window=fixed_window, lib=lib)
    assert ver.verify([pre + tu for tu, _ in got], [pr for _, pr in got]) == [True] * B
    ver.close()


def check_prove_with_identity_trackers(lib, name, fixed_window=0, table_window=None):
    """Identity points really occur on the wire.  Pre-shuffle trackers that ARE the identity give identity
    post-shuffle trackers, all-infinity rows in the prover's tables of multiples and infinities inside its MSMs: the
    batched prover must still emit the bytes the oracle's per-proof restatement emits under the same randomness, and
    the batched verifier must accept them."""
    from oracle import ark_surface
    from oracle.shuffle_ref import ShuffleRef

    ark_surface.set_backend("c")
    G1Point, Scalar = ark_surface.G1Point, ark_surface.Scalar
    case = sc.load_case(name)
    ell = case["N"] - 4
    crs_bytes = bytes.fromhex(case["crs"])
    INF = bytes([0xC0]) + bytes(47)
    R_hex, S_hex = list(case["vec_R"]), list(case["vec_S"])
    R_hex[1] = INF.hex(); S_hex[1] = INF.hex(); S_hex[ell - 1] = INF.hex()
    if ell >= 4:                                         # and a repeated tracker: equal points meet inside the MSMs
        R_hex[3], S_hex[3] = R_hex[2], S_hex[2]
    pre = b"".join(bytes.fromhex(h) for h in R_hex + S_hex)
    # oracle: reference draw order perm -> k -> m_bl(4) -> CurdleProofsProof.new's draws
    ctx = ShuffleRef(G1Point, Scalar, rng=random.Random(4711))
    crs = ctx.crs_from_bytes(crs_bytes, ell)
    dec = lambda lst: [G1Point.from_compressed_bytes_unchecked(bytes.fromhex(h)) for h in lst]  # noqa: E731
    vec_R, vec_S = dec(R_hex), dec(S_hex)
    perm = list(range(ell)); ctx.rng.shuffle(perm)
    k = ctx.rand()
    vec_T, vec_U, M, m_bl = ctx.shuffle_and_commit(crs, vec_R, vec_S, perm, k)
    want_proof = ctx.prove(crs, vec_R, vec_S, vec_T, vec_U, M, perm, k, m_bl)
    want_tu = b"".join(ctx.pb(p) for p in vec_T + vec_U)
    # the batched prover on the same stream of draws
    prover = whisk.BatchProver(crs_bytes, ell, fixed_window=fixed_window, lib=lib)
    if table_window is not None:
        prover.set_table_window(table_window)
    (tu, proof), = prover.prove_drawn([pre], random.Random(4711))
    prover.close()
    assert tu == want_tu, "post-shuffle trackers differ from the oracle's"
    assert INF in (tu[48 * i:48 * i + 48] for i in range(2 * ell))
    assert proof == ctx.pb(M) + want_proof, "proof bytes differ from the oracle's"
    ver = whisk.BatchVerifier(crs_bytes, ell, fixed_window=fixed_window, lib=lib)
    assert ver.verify([pre + tu], [proof]) == [ctx.is_valid(crs, vec_R, vec_S, vec_T, vec_U, M, want_proof)]
    ver.close()


def check_rejects_bad_perm_and_blinders(lib, name, fixed_window=0):
    """perm entries index device rows - an entry >= ell, a repeated entry or a blinder >= r must flag that
    lane (status 1) without touching the device with it, while the clean lane still yields the reference's bytes; the
    Python wrappers refuse wrong lengths before libcpg reads past a buffer."""
    import pytest

    case = sc.load_case(name)
    ell = case["N"] - 4
    prover = whisk.BatchProver(bytes.fromhex(case["crs"]), ell, fixed_window=fixed_window, lib=lib)
    perm, k, rand = replay_rng(case, prover)
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    oob = list(perm); oob[1] = ell + 7                       # out of range
    dup = list(perm); dup[0] = dup[1]                        # not a permutation
    huge = list(perm); huge[2] = 0xFFFFFFFF
    bad_rand = bytearray(rand); bad_rand[32 * 5:32 * 6] = b"\xff" * 32   # blinder >= r
    kb = k.to_bytes(32, "little")
    tu, pr, st = prover.prove_raw(pre * 5, perm + oob + dup + huge + perm, kb * 5, rand * 4 + bytes(bad_rand), 5)
    assert list(st) == [0, 1, 1, 1, 1]
    assert pr[:prover.proof_len] == bytes.fromhex(case["M"]) + bytes.fromhex(case["proof"])
    with pytest.raises(IndexError):
        prover.prove([pre], [oob], [k], [rand])
    with pytest.raises(ValueError):
        prover.prove([pre], [dup], [k], [rand])              # flagged by the library -> ValueError
    with pytest.raises(ValueError):
        prover.prove([pre[:-48]], [perm], [k], [rand])       # short tracker row
    with pytest.raises(ValueError):
        prover.prove([pre], [perm[:-1]], [k], [rand])
    with pytest.raises(ValueError):
        prover.prove([pre], [perm], [k], [rand[:-32]])
    with pytest.raises(ValueError):
        prover.prove_raw(pre, perm, kb, rand[:-1], 1)
    with pytest.raises(ValueError):
        prover.prove_drawn([pre, pre[:-1]], random.Random(1))
    prover.close()


def check_whisk_api_roundtrip(lib, name, B=3):
    """GenerateWhiskShuffleProofBatch -> IsValidWhiskShuffleProofBatch (the batched forms of whisk_interface.py:74-140):
    trackers in, WhiskTracker-shaped objects + proof bytes out, every proof valid, and the first proof equals what the
    reference's per-proof path draws from the same `random` stream (replayed through prove())."""
    import pytest

    case = sc.load_case(name)
    ell = case["N"] - 4
    crs = (bytes.fromhex(case["crs"]), ell)
    pre = [(bytes.fromhex(r), bytes.fromhex(s)) for r, s in zip(case["vec_R"], case["vec_S"])]
    whisk._CACHE.clear(); whisk._PROVER_CACHE.clear()
    import curdleproofs_pie_b200.runtime as _rt
    old = _rt._LIB if hasattr(_rt, "_LIB") else None
    try:
        _rt._install_library_for_tests(lib)
        a, b = random.Random(8128), random.Random(8128)
        res = whisk.GenerateWhiskShuffleProofBatch(crs, [pre] * B, rng=a)
        assert len(res) == B
        post0, proof0 = res[0]
        assert all(isinstance(t, whisk.WhiskTracker) and len(t.r_G) == 48 and len(t.k_r_G) == 48 for t in post0)
        assert whisk.IsValidWhiskShuffleProofBatch(crs, [pre] * B, [post for post, _ in res], [pr for _, pr in res]) == [True] * B
        # tuples are accepted on the way back in, and a proof paired with another lane's trackers is rejected
        as_tuples = [[tuple(t) for t in post] for post, _ in res]
        assert whisk.IsValidWhiskShuffleProofBatch(crs, [pre] * B, as_tuples[1:] + as_tuples[:1], [pr for _, pr in res]) == [False] * B
        # same draws through the explicit-randomness entry point
        prover = whisk._PROVER_CACHE[(crs[0], ell, 4)]
        p = list(range(ell)); b.shuffle(p)
        k = b.randint(1, rt.R_ORDER - 1)
        (tu, pr), = prover.prove([b"".join(r for r, _ in pre) + b"".join(s_ for _, s_ in pre)], [p], [k], [prover.draw_randomness(b)])
        assert pr == proof0 and tu == b"".join(t.r_G for t in post0) + b"".join(t.k_r_G for t in post0)
        with pytest.raises(ValueError):
            whisk.GenerateWhiskShuffleProofBatch(crs, [pre[:-1]], rng=a)
        with pytest.raises(ValueError):
            whisk.GenerateWhiskShuffleProofBatch(crs, [pre[:-1] + [(pre[0][0], pre[0][1][:-1])]], rng=a)
    finally:
        for c_ in (whisk._CACHE, whisk._PROVER_CACHE):
            for obj in c_.values():
                obj.close()
            c_.clear()
        _rt._install_library_for_tests(old)

"""Batched verifier (cpg_verify_batch) against the golden fixtures produced by the UNMODIFIED
reference: per-proof verdicts must equal the reference's on honest and corrupted inputs."""
import shuffle_cases as sc
from curdleproofs_pie_b200 import whisk


def _inputs(case, R=None, S=None, T=None, U=None):
    g = lambda k, o: b"".join(bytes.fromhex(h) for h in (o if o is not None else case[k]))  # noqa: E731
    return g("vec_R", R) + g("vec_S", S) + g("vec_T", T) + g("vec_U", U)


def variants(case):
    """(name, inputs, proof, expected) for the fixture's corrupted variants plus byte-level ones."""
    M, proof = bytes.fromhex(case["M"]), bytes.fromhex(case["proof"])
    good = _inputs(case)
    v = case["verdicts"]
    out = [
        ("honest", good, M + proof, v["honest"]),
        ("swap_R_S", _inputs(case, R=case["vec_S"], S=case["vec_R"]), M + proof, v["swap_R_S"]),
        ("swap_T_U", _inputs(case, T=case["vec_U"], U=case["vec_T"]), M + proof, v["swap_T_U"]),
        ("rotated_T", _inputs(case, T=case["vec_T"][1:] + case["vec_T"][:1]), M + proof, v["rotated_T"]),
    ]
    # a flipped bit in every scalar / a few points of the proof must be rejected (AssertionError or
    # ValueError in the reference, both -> False at the Whisk boundary)
    lg = case["N"].bit_length() - 1
    scalar_offsets = [48 * 9, 48 * 9 + 32 + 48 * (2 + 4 * lg), 48 * 9 + 32 + 48 * (2 + 4 * lg) + 64 + 48 * 4]
    for off in scalar_offsets + [0, 48 * 7 + 5, len(proof) - 30]:
        bad = bytearray(proof)
        bad[off + 1] ^= 0x04
        out.append(("flip@%d" % off, good, M + bytes(bad), False))
    # every scalar of the proof, and one point of every named field / vector replaced by ANOTHER VALID point (the CRS's
    # H): such a proof still decodes, so only the coefficient that verify_phase2 gives this very term can reject it
    n = case["N"]
    crs = bytes.fromhex(case["crs"])
    Hb = crs[48 * n:48 * n + 48]
    o_rp = 48 * 9
    o_ipa = o_rp + 32 + 48 * 2                     # L_C | R_C | L_D | R_D, lg points each
    o_cd = o_ipa + 48 * 4 * lg
    o_cm = o_cd + 64                               # cm_A (2) | cm_B (2)
    o_z = o_cm + 48 * 4
    o_b = o_z + 96                                 # B_a B_t B_u
    o_msm = o_b + 48 * 3                           # L_A L_T L_U R_A R_T R_U, lg points each
    o_x = o_msm + 48 * 6 * lg
    assert o_x + 32 == len(proof)
    for nm, off in [("r_p", o_rp), ("c", o_cd), ("d", o_cd + 32), ("z_k", o_z), ("z_t", o_z + 32), ("z_u", o_z + 64), ("x", o_x)]:
        bad = bytearray(proof)
        bad[off] ^= 0x01
        out.append(("scalar:" + nm, good, M + bytes(bad), False))
    named = [("A", 0), ("T1", 48), ("T2", 96), ("U1", 144), ("U2", 192), ("R", 240), ("S", 288), ("B", 336), ("C", 384),
             ("Bc", o_rp + 32), ("Bd", o_rp + 32 + 48), ("cmA1", o_cm), ("cmA2", o_cm + 48), ("cmB1", o_cm + 96), ("cmB2", o_cm + 144),
             ("Ba", o_b), ("Bt", o_b + 48), ("Bu", o_b + 96)]
    for k, nm in enumerate(["L_C", "R_C", "L_D", "R_D"]):
        named.append((nm + "[0]", o_ipa + 48 * k * lg))
        named.append((nm + "[last]", o_ipa + 48 * (k * lg + lg - 1)))
    for k, nm in enumerate(["L_A", "L_T", "L_U", "R_A", "R_T", "R_U"]):
        named.append((nm + "[0]", o_msm + 48 * k * lg))
        named.append((nm + "[last]", o_msm + 48 * (k * lg + lg - 1)))
    for nm, off in named:
        bad = bytearray(proof)
        assert bytes(bad[off:off + 48]) != Hb
        bad[off:off + 48] = Hb
        out.append(("point:" + nm, good, M + bytes(bad), False))
    out.append(("point:M", good, Hb + proof, False))
    non_canonical = bytearray(proof)
    non_canonical[48 * 9:48 * 9 + 32] = b"\xff" * 32            # r_p >= r: from_le_bytes raises
    out.append(("scalar>=r", good, M + bytes(non_canonical), False))
    inf_T0 = bytearray(good)
    ell = case["N"] - 4
    inf_T0[48 * 2 * ell:48 * 2 * ell + 48] = bytes([0xC0]) + bytes(47)   # vec_T[0] = identity -> Exception
    out.append(("T0=inf", bytes(inf_T0), M + proof, False))
    return out


def check_batch(lib, name, copies=1, window=0, transcript_on_device=True, fixed_window=0, group=1, sharded=False):
    case = sc.load_case(name)
    ell = case["N"] - 4
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), ell, fixed_window=fixed_window, lib=lib, sharded=sharded)
    ver.set_transcript(transcript_on_device)
    ver.set_group(group)
    if window:
        ver.set_window(window)
    vs = variants(case) * copies
    got = ver.verify([v[1] for v in vs], [v[2] for v in vs])
    want = [v[3] for v in vs]
    assert got == want, [(v[0], g, w) for v, g, w in zip(vs, got, want) if g != w]
    # wrong lengths are rejected without touching the device result of the others
    got = ver.verify([vs[0][1], vs[0][1][:-1], vs[0][1]], [vs[0][2], vs[0][2], vs[0][2][:-1]])
    assert got == [True, False, False]
    if group > 1:
        # all-honest groups are settled by the aggregated check alone; one bad lane sends only its group back
        honest = vs[0]
        nb = 3 * group + 1
        assert ver.verify([honest[1]] * nb, [honest[2]] * nb) == [True] * nb
        assert ver.rechecked() == 0
        bad = vs[1]
        ins = [honest[1]] * nb; prs = [honest[2]] * nb
        ins[group + 1], prs[group + 1] = bad[1], bad[2]
        assert ver.verify(ins, prs) == [i != group + 1 for i in range(nb)]
        assert ver.rechecked() == group
    ver.close()


def check_cache(lib, name, log2_slots=12, transcript_on_device=True, group=1, streams=0, copies=3):
    """The decompressed-tracker cache (cpg_verifier_set_cache) must not change a single verdict: the fixture's variants
    (which share most trackers: duplicates inside the batch), a malformed tracker encoding (cached with its error code),
    then the same batch again (served from the table), then a table so small that it overflows and starts over."""
    case = sc.load_case(name)
    ell = case["N"] - 4
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), ell, lib=lib)
    ver.set_transcript(transcript_on_device)
    ver.set_group(group)
    if streams:
        ver.set_streams(streams)
    vs = variants(case)
    bad_enc = bytearray(vs[0][1]); bad_enc[48 * 3] &= 0x7F                      # vec_R[3]: compression flag missing
    vs.append(("tracker:bad-encoding", bytes(bad_enc), vs[0][2], False))
    vs = vs * copies
    want = [v[3] for v in vs]
    ins, prs = [v[1] for v in vs], [v[2] for v in vs]
    assert ver.verify(ins, prs) == want                                          # cache off
    ver.set_cache(log2_slots)
    got = ver.verify(ins, prs)
    assert got == want, [(v[0], g, w) for v, g, w in zip(vs, got, want) if g != w]
    st1 = ver.cache_stats()
    assert st1["lookups"] == len(vs) * 4 * ell and 0 < st1["claimed"] <= 5 * ell + 8, st1     # 4 ell distinct trackers + the few altered ones
    # sub-batches on concurrent streams may meet a slot that another one is still filling: they decompress that point
    # themselves (uncached) instead of waiting, so with several streams `served` is only bounded from above
    if streams <= 1 or len(vs) < 2048:                       # (a sub-batch holds at least 1024 proofs)
        assert st1["served"] == st1["lookups"] - st1["claimed"], st1
    else:
        assert 0 < st1["served"] <= st1["lookups"] - st1["claimed"], st1
    got = ver.verify(ins, prs)
    assert got == want
    st2 = ver.cache_stats()
    assert st2["claimed"] == st1["claimed"] and st2["served"] - st1["served"] == len(vs) * 4 * ell, (st1, st2)   # all from the table
    ver.cache_reset()
    assert ver.cache_stats() == {"lookups": 0, "served": 0, "claimed": 0}
    assert ver.verify(ins, prs) == want
    # a table of 1024 slots with 4 ell = 496+ distinct keys per batch: more than half full after one batch -> it starts
    # over before the next one, and probe sequences that run out fall back to plain decompression
    ver.set_cache(10)
    for _ in range(3):
        assert ver.verify(ins, prs) == want
    ver.set_cache(0)
    assert ver.verify(ins, prs) == want
    ver.close()

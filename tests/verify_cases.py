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
    non_canonical = bytearray(proof)
    non_canonical[48 * 9:48 * 9 + 32] = b"\xff" * 32            # r_p >= r: from_le_bytes raises
    out.append(("scalar>=r", good, M + bytes(non_canonical), False))
    inf_T0 = bytearray(good)
    ell = case["N"] - 4
    inf_T0[48 * 2 * ell:48 * 2 * ell + 48] = bytes([0xC0]) + bytes(47)   # vec_T[0] = identity -> Exception
    out.append(("T0=inf", bytes(inf_T0), M + proof, False))
    return out


def check_batch(lib, name, copies=1, window=0, transcript_on_device=True, fixed_window=0, group=1):
    case = sc.load_case(name)
    ell = case["N"] - 4
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), ell, fixed_window=fixed_window, lib=lib)
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

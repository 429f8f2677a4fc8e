#!/usr/bin/env python
"""The UNMODIFIED reference (baseline/_ref: curdleproofs.CurdleProofsProof.new / .verify and the Whisk API,
/root/reference/curdleproofs/curdleproofs/curdleproofs.py:50-248, whisk_interface.py:74-140) running on the B200 drop-in:
`py_arkworks_bls12381` = dropin/py_arkworks_bls12381 (libcpg.so kernels), `merlin_transcripts` = dropin/merlin_transcripts
(libcpg.so's STROBE/Keccak).  Replays a golden fixture's construction under its seed, checks that the reference's own
code then emits the fixture's bytes and verdicts, and times it.  Prints ONE JSON line.

    python tools/reference_on_dropin.py [--case shuffle_N128_seed4096.json] [--repeat 2] [--python-merlin]

Run as its own process: which `py_arkworks_bls12381` a process sees is decided by sys.path at first import.
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="shuffle_N128_seed4096.json")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--python-merlin", action="store_true", help="keep the reference's own pure-Python merlin_transcripts (0.7 ms per Keccak-f)")
    ap.add_argument("--stats", action="store_true", help="also report where the time of .new / .verify goes: count and seconds of every observation (MSM by size class), decompression, compression and transcript call of the drop-ins")
    ap.add_argument("--defer-decode", action="store_true", help="drop-in: deferred decoding (from_compressed_bytes* only records the bytes; one batched decode at first use; a malformed encoding raises at its first use instead)")
    ap.add_argument("--test-seam", action="store_true", help="CPU test tier only: run on the host emulation of the kernels (tests/conftest.py::build_seam)")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "curdleproofs")):
        print(json.dumps({"unavailable": "baseline/_ref is not installed (python tools/install_reference.py, needs /root/reference)"}))
        return
    dropin = os.path.join(ROOT, "dropin")
    if args.test_seam:
        sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
        import conftest
        from curdleproofs_pie_b200 import runtime as rt0

        rt0._install_library_for_tests(rt0.CpgLib(conftest.build_seam(), 0))
    if args.python_merlin:
        # only the arithmetic drop-in: expose dropin/py_arkworks_bls12381 without dropin/merlin_transcripts
        import importlib.util

        sys.path[:0] = [REF, ROOT]
        spec = importlib.util.spec_from_file_location("py_arkworks_bls12381", os.path.join(dropin, "py_arkworks_bls12381", "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["py_arkworks_bls12381"] = mod
        spec.loader.exec_module(mod)
    else:
        sys.path[:0] = [dropin, REF, ROOT]
    import merlin_transcripts
    import py_arkworks_bls12381 as ark
    from curdleproofs_pie_b200 import runtime as rt

    lib = rt.get_lib()
    ark.defer_decoding(args.defer_decode)
    assert lib.backend == ("host-emulation-test-seam" if args.test_seam else "cuda-sm_100a"), lib.backend
    assert os.path.dirname(ark.__file__).startswith(dropin), ark.__file__
    import curdleproofs
    from curdleproofs.crs import CurdleproofsCrs
    from curdleproofs.curdleproofs import N_BLINDERS, CurdleProofsProof, shuffle_permute_and_commit_input
    from curdleproofs.util import BufReader, get_random_point, point_projective_to_bytes, random_scalar
    from curdleproofs.whisk_interface import GenerateWhiskShuffleProof, IsValidWhiskShuffleProof, WhiskTracker

    assert os.path.dirname(curdleproofs.__file__).startswith(REF), curdleproofs.__file__
    with open(os.path.join(ROOT, "tests", "golden", args.case)) as f:
        case = json.load(f)
    N = case["N"]
    ell = N - N_BLINDERS
    stats, phase = {}, ["setup"]
    if args.stats:
        def timed(owner, name, label, static=False):
            fn = getattr(owner, name)

            def wrap(*a, **kw):
                lab = label(*a) if callable(label) else label
                t0 = time.perf_counter()
                try:
                    return fn(*a, **kw)
                finally:
                    e = stats.setdefault(phase[0], {}).setdefault(lab, [0, 0.0])
                    e[0] += 1
                    e[1] += time.perf_counter() - t0
            setattr(owner, name, staticmethod(wrap) if static else wrap)

        def force_label(p):
            if p._aff is not None:
                return "force(cached)"
            if getattr(p, "_dec", None) is not None:
                return "force (flush of the recorded decodes)"
            n = len(p._terms)
            return "force n=%s" % (n if n <= 2 else "3-8" if n <= 8 else "9-64" if n <= 64 else "65-256" if n <= 256 else ">256")
        timed(ark.G1Point, "_force", force_label)
        timed(ark.G1Point, "_decompress", "decompress", static=True)
        timed(ark.G1Point, "to_compressed_bytes", "to_compressed_bytes (incl. force)")
        timed(merlin_transcripts.MerlinTranscript, "append_message", "merlin.append_message")
        timed(merlin_transcripts.MerlinTranscript, "challenge_bytes", "merlin.challenge_bytes")
    launches0 = lib.launch_count()
    t_new, t_verify, t_verify_warm, t_whisk_v, t_whisk_p, t_whisk_v_cold = [], [], [], [], [], []
    for rep in range(args.repeat):
        # the fixture's construction order (oracle/gen_golden.py; cp/test_curdleproofs.py:576-593)
        random.seed(case["seed"])
        crs = CurdleproofsCrs.new(ell, N_BLINDERS)
        perm = list(range(ell))
        random.shuffle(perm)
        k = random_scalar()
        vec_R = [get_random_point() for _ in range(ell)]
        vec_S = [get_random_point() for _ in range(ell)]
        vec_T, vec_U, M, bl = shuffle_permute_and_commit_input(crs, vec_R, vec_S, perm, k)
        phase[0] = "new"
        t0 = time.perf_counter()
        proof = CurdleProofsProof.new(crs=crs, vec_R=vec_R, vec_S=vec_S, vec_T=vec_T, vec_U=vec_U, M=M, permutation=perm, k=k, vec_m_blinders=bl)
        wire = proof.to_bytes()
        t_new.append(time.perf_counter() - t0)
        phase[0] = "other"
        assert crs.to_bytes().hex() == case["crs"], "CRS bytes differ from the fixture"
        enc = lambda pts: [point_projective_to_bytes(p).hex() for p in pts]  # noqa: E731
        assert enc(vec_T) == case["vec_T"] and enc(vec_U) == case["vec_U"] and point_projective_to_bytes(M).hex() == case["M"], "shuffle outputs differ"
        assert wire.hex() == case["proof"], "proof bytes differ from the fixture written by the reference on the CPU oracle"

        def verdict(R_, S_, T_, U_, M_):
            try:
                CurdleProofsProof.from_bytes(BufReader(wire), N).verify(crs, R_, S_, T_, U_, M_)
                return True
            except AssertionError:
                return False

        # cold: no encoding of this proof is in the drop-in's decode cache (a verifier that never saw these bytes);
        # warm: the same call again (the inputs and the proof were decoded once already)
        getattr(ark, "_decoded", {}).clear()
        phase[0] = "verify"
        t0 = time.perf_counter()
        honest = verdict(vec_R, vec_S, vec_T, vec_U, M)
        t_verify.append(time.perf_counter() - t0)
        phase[0] = "verify_warm"
        t0 = time.perf_counter()
        assert verdict(vec_R, vec_S, vec_T, vec_U, M) == honest
        t_verify_warm.append(time.perf_counter() - t0)
        phase[0] = "other"
        got = {"honest": honest, "swap_R_S": verdict(vec_S, vec_R, vec_T, vec_U, M), "swap_T_U": verdict(vec_R, vec_S, vec_U, vec_T, M),
               "wrong_M": verdict(vec_R, vec_S, vec_T, vec_U, M + M), "rotated_T": verdict(vec_R, vec_S, vec_T[1:] + vec_T[:1], vec_U, M)}
        assert got == case["verdicts"], (got, case["verdicts"])
        # the Whisk-facing API, bytes in / bytes out (whisk_interface.py:74-140)
        pre = [WhiskTracker(bytes.fromhex(r), bytes.fromhex(s)) for r, s in zip(case["vec_R"], case["vec_S"])]
        post = [WhiskTracker(bytes.fromhex(t), bytes.fromhex(u)) for t, u in zip(case["vec_T"], case["vec_U"])]
        whisk_wire = bytes.fromhex(case["M"]) + wire
        getattr(ark, "_decoded", {}).clear()
        t0 = time.perf_counter()
        ok = IsValidWhiskShuffleProof(crs, pre, post, whisk_wire)
        t_whisk_v_cold.append(time.perf_counter() - t0)
        assert ok is True
        t0 = time.perf_counter()
        ok = IsValidWhiskShuffleProof(crs, pre, post, whisk_wire)
        t_whisk_v.append(time.perf_counter() - t0)
        assert ok is True
        assert IsValidWhiskShuffleProof(crs, post, pre, whisk_wire) is False
        t0 = time.perf_counter()
        post2, wire2 = GenerateWhiskShuffleProof(crs, pre)
        t_whisk_p.append(time.perf_counter() - t0)
        assert IsValidWhiskShuffleProof(crs, pre, post2, wire2) is True
    lib.sync()
    extra = {"stats": {ph: {k: [v[0], round(v[1], 4)] for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])} for ph, d in stats.items()}} if args.stats else {}
    print(json.dumps({**extra, 
        "what": "the UNMODIFIED reference (baseline/_ref) on the B200 drop-in; proof bytes and 5 verdicts equal tests/golden/%s" % args.case,
        "n": N, "merlin": "reference pure-Python" if args.python_merlin else "dropin/merlin_transcripts (libcpg.so STROBE/Keccak)",
        "merlin_module": os.path.relpath(merlin_transcripts.__file__, ROOT),
        "CurdleProofsProof_new_s": min(t_new), "CurdleProofsProof_verify_s": min(t_verify), "CurdleProofsProof_verify_warm_s": min(t_verify_warm),
        "IsValidWhiskShuffleProof_s": min(t_whisk_v_cold), "IsValidWhiskShuffleProof_warm_s": min(t_whisk_v),
        "cache_note": "verify / IsValid: the drop-in's encoding -> point map emptied before the call (all 4 ell trackers and the proof's points are decompressed on the GPU); _warm: the same call again", "GenerateWhiskShuffleProof_s": min(t_whisk_p),
        "deferred_decoding": bool(args.defer_decode), "repeat": args.repeat, "gpu_launches": lib.launch_count() - launches0, "backend": lib.backend}))


if __name__ == "__main__":
    main()

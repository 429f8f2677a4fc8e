"""BASELINE config 5 on ONE GPU: a single n = 16384 shuffle proof through cpg_prove_batch / cpg_verify_batch,
checked against the digests of the unmodified reference's outputs (tests/golden/large_N16384_seed21384.json).
    python tools/large_proof.py [fixture.json] > gpurun_out/large_proof.json
The batched pipeline is laid out for thousands of proofs (one thread per proof in the transcript / Fr kernels),
so a lone proof is latency-bound; this script records that a proof of this size runs and is bit-exact."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import __graft_entry__ as ge

    ge.build()
    import large_cases as lc
    from curdleproofs_pie_b200 import runtime as rt

    name = sys.argv[1] if len(sys.argv) > 1 else "large_N16384_seed21384.json"
    lib = rt.get_lib()
    lib.profile(True)
    t0 = time.perf_counter()
    res = lc.check_large(lib, name, fixed_window=int(os.environ.get("CPG_LARGE_FIXED_WINDOW", "8")))
    res["total_s"] = time.perf_counter() - t0
    prof = lib.profile_report()
    lib.profile(False)
    res["kernels_ms"] = {k: round(v["ms"], 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]}
    res["fixture"] = name
    res["parity"] = "post-shuffle trackers, M, proof bytes (SHA-256) and 3 verdicts equal the unmodified reference's"
    print(json.dumps(res))


if __name__ == "__main__":
    main()

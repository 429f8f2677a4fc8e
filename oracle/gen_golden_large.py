"""ORACLE tooling: compact golden fixtures for LARGE shuffles (BASELINE config 5: n = 16384, plus n = 1024
as a test-sized stand-in), produced like oracle/gen_golden.py by running the UNMODIFIED reference on the
oracle's arithmetic under a fixed seed.  The inputs of such a case are megabytes of points, so the fixture
keeps only the seed and SHA-256 digests; a test re-creates the inputs from the seed (every point is
s_i * G with s_i drawn from Python's `random` in the reference's order) and compares digests.

    python oracle/gen_golden_large.py 1024 16384          # needs /root/reference; n = 16384 takes minutes
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden as gg  # noqa: E402  (sets up sys.path for the reference and the stand-in)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    gg.ref_strobe.KeccakF1600 = gg._fast_f1600
    for N in [int(a) for a in sys.argv[1:]] or [1024]:
        t0 = time.time()
        seed = 5000 + N
        case = gg.one_case(seed, N)
        cat = lambda k: b"".join(bytes.fromhex(h) for h in case[k])  # noqa: E731
        compact = {
            "seed": seed, "N": N, "keccak": "c", "k": case["k"], "perm_sha256": sha(json.dumps(case["perm"]).encode()),
            "crs_sha256": sha(bytes.fromhex(case["crs"])), "pre_sha256": sha(cat("vec_R") + cat("vec_S")),
            "post_sha256": sha(cat("vec_T") + cat("vec_U")), "M": case["M"],
            "proof_len": len(case["proof"]) // 2, "proof_sha256": case["proof_sha256"], "verdicts": case["verdicts"],
        }
        path = os.path.join(out_dir, "large_N%d_seed%d.json" % (N, seed))
        with open(path, "w") as f:
            json.dump(compact, f, indent=1)
        print(path, compact["proof_len"], "bytes", compact["proof_sha256"][:16], compact["verdicts"], "%.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()

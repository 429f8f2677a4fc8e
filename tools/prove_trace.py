#!/usr/bin/env python
"""Timeline of ONE cpg_prove_batch call with B proofs (default 1): every kernel launch with start / end in ms
(CPG_PROFILE_TRACE) next to the wall clock of the call.
    python tools/prove_trace.py 1 gpurun_out/prove_trace_B1.txt"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prove_trace.txt"
os.environ["CPG_PROFILE_TRACE"] = path
import bench  # noqa: E402
from curdleproofs_pie_b200 import runtime as rt  # noqa: E402
from curdleproofs_pie_b200 import whisk  # noqa: E402

lib = rt.get_lib()
case = bench.load_golden()
crs = bytes.fromhex(case["crs"])
cat = lambda k: b"".join(bytes.fromhex(h) for h in case[k])  # noqa: E731
pre = cat("vec_R") + cat("vec_S")
prover = whisk.BatchProver(crs, bench.ELL, fixed_window=12)
ver = whisk.BatchVerifier(crs, bench.ELL, fixed_window=12)
rng = random.Random(case["seed"])
ts = []
for _ in range(7):
    t0 = time.perf_counter()
    res = prover.prove_drawn([pre] * B, rng)
    ts.append((time.perf_counter() - t0) * 1e3)
assert ver.verify([pre + res[0][0]], [res[0][1]]) == [True]
lib.check(lib.c.cpg_profile_reset()); lib.profile(True)
t0 = time.perf_counter()
res = prover.prove_drawn([pre] * B, rng)
t_prof = (time.perf_counter() - t0) * 1e3
rep = lib.profile_report()
lib.profile(False)
print("B = %d: median wall clock %.3f ms per call (min %.3f); with per-launch events %.3f ms; %d launches; kernel time %.3f ms; timeline in %s"
      % (B, sorted(ts)[len(ts) // 2], min(ts), t_prof, sum(v["launches"] for v in rep.values()), sum(v["ms"] for v in rep.values()), path))
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:16]:
    print("  %-20s %7.3f ms  %d launch(es)" % (k, v["ms"], v["launches"]))

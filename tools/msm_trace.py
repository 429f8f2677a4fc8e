#!/usr/bin/env python
"""Timeline of ONE n-term MSM (cpg_g1_msm_batched, B = 1): every kernel launch with its start and end in ms (CUDA event
timestamps, comparable across streams), to see what the pipelined pipeline (CPG_MSM_SLICES) overlaps.
    CPG_MSM_SLICES=4 python tools/msm_trace.py 20 gpurun_out/trace.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/msm_trace.txt"
os.environ["CPG_PROFILE_TRACE"] = path
import bench  # noqa: E402
from curdleproofs_pie_b200 import runtime as rt  # noqa: E402

lib = rt.get_lib()
n = 1 << lg
uaff, nuniq, bases, scalars, sc_bytes = bench.large_msm_instance(lib, rt, n)
c = int(lib.c.cpg_msm_pick_window(n))
out = lib.alloc(rt.JAC)
for _ in range(4):
    lib.check(lib.c.cpg_g1_msm_batched(bases.ptr, 0, scalars.ptr, 1, n, c, out.ptr), "msm")
lib.sync()
want = bench.oracle_large_msm(lib, uaff, nuniq, sc_bytes, n)
assert lib.compress_jac(out, 1) == want, "MSM result differs from the oracle"
lib.check(lib.c.cpg_profile_reset()); lib.profile(True)
lib.timer_start()
lib.check(lib.c.cpg_g1_msm_batched(bases.ptr, 0, scalars.ptr, 1, n, c, out.ptr), "msm")
ms = lib.timer_stop()
lib.profile_report()
lib.profile(False)
print("n = 2^%d, window %d, slices %s: %.3f ms (with per-launch events); timeline in %s" % (lg, c, os.environ.get("CPG_MSM_SLICES", "default"), ms, path))

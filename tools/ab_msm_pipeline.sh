#!/bin/bash
# A/B of the pipelined single large MSM (CPG_MSM_SLICES window slices; 1 = the plain pipeline), forced on for every size
for sl in 1 2 4 8; do
  CPG_MSM_SLICES=$sl CPG_MSM_PIPE_MIN_N=1 python bench.py --workload msm_sweep --steps 3 --warmup 3 > gpurun_out/r02_ab_msm_slices$sl.json 2> gpurun_out/r02_ab_msm_slices$sl.err
  python - <<P
import json
d=json.load(open("gpurun_out/r02_ab_msm_slices$sl.json"))
print("slices=$sl", " ".join("2^%d:%.2fms" % (s["n"].bit_length()-1, s["ms"]) for s in d["sizes"] if s["n"] >= 1<<15))
P
done

#!/bin/bash
# A/B of the resident blocks per SM of the unchecked Decompress kernel (CPG_DECOMP_MINB: 3 = the other point kernels'
# launch, 5 = default), verify workload only
for mb in 3 5; do
  CPG_DECOMP_MINB=$mb python bench.py --workload verify --steps 3 --warmup 3 > gpurun_out/r02_ab_decomp_minb$mb.json 2> gpurun_out/r02_ab_decomp_minb$mb.err
  python - <<P
import json
d=json.load(open("gpurun_out/r02_ab_decomp_minb$mb.json"))
k=d["roofline"]["kernels"]
print("minb=$mb value=%.0f e2e=%.0f ms_step=%.2f Decompress=%.2f ms frac=%.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], k["Decompress"]["ms_per_step"], k["Decompress"]["frac"]))
P
done

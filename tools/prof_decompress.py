#!/usr/bin/env python
"""Three launches of the unchecked Decompress kernel over 2^21 valid encodings (4096 distinct points s_i G, tiled), for
`ncu --set full -k regex:Decompress` (tools/ncu_traffic.py --units Decompress=2097152 reads the capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from curdleproofs_pie_b200 import runtime as rt  # noqa: E402

lib = rt.get_lib()
K, N = 4096, 1 << 21
gens = lib.upload(lib.download(lib.generator(), rt.JAC) * K)
scalars = lib.upload(b"".join((0x9E3779B97F4A7C15 * (i + 1) % rt.R_ORDER).to_bytes(32, "little") for i in range(K)))
enc = lib.compress_jac(lib.mul(gens, scalars, K), K)
data = enc * (N // K)
for _ in range(3):
    out, err = lib.decompress(data)
    assert not any(err)
lib.sync()
print("decompressed", N, "points x 3; backend", lib.backend)

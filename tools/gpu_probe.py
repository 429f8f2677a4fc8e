"""GPU probe: integer-pipe rates and first timings of every kernel family (not a bench line)."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dropin")]
from curdleproofs_pie_b200 import runtime as rt  # noqa: E402

lib = rt.get_lib()
out = {"backend": lib.backend}


def timed(fn, reps=3):
    fn(); lib.sync()
    best = 1e30
    for _ in range(reps):
        lib.timer_start(); fn(); best = min(best, lib.timer_stop())
    return best


for kind, name in ((0, "imad_wide_mac_per_s"), (1, "imad_lohi_mac_per_s"), (2, "fq_mul_per_s")):
    per_s, ms = lib.bench_int_pipe(kind, 20000 if kind < 2 else 2000)
    out[name] = per_s
    print(name, "%.4g" % per_s, "ms", ms, flush=True)

rng = random.Random(5)


def random_points(k):
    gen = lib.generator()
    gens = lib.alloc(k * rt.JAC)
    # replicate the generator by doubling copies
    lib.check(lib.c.cpg_d2d(gens.ptr, gen.ptr, rt.JAC))
    have = 1
    while have < k:
        cnt = min(have, k - have)
        lib.check(lib.c.cpg_d2d(gens.ptr + have * rt.JAC, gens.ptr, cnt * rt.JAC))
        have += cnt
    ks = lib.upload(os.urandom(31 * k).join([b""]) if False else b"".join(rng.randrange(rt.R_ORDER).to_bytes(32, "little") for _ in range(k)))
    return lib.mul(gens, ks, k)


def rand_scalars(k):
    return lib.upload(b"".join(rng.randrange(rt.R_ORDER).to_bytes(32, "little") for _ in range(k)))


K = 1 << 17
t0 = time.time()
jac = random_points(K)
lib.sync()
print("setup random points", K, "in", time.time() - t0, "s", flush=True)
ks = rand_scalars(K)
ms = timed(lambda: lib.mul(jac, ks, K))
out["mul_points_per_s"] = K / ms * 1e3
print("elementwise mul: %.1f ms for %d -> %.3g /s" % (ms, K, K / ms * 1e3), flush=True)
aff = lib.jac_to_aff(jac, K)
ms = timed(lambda: lib.jac_to_aff(jac, K))
print("jac_to_aff: %.2f ms -> %.3g /s" % (ms, K / ms * 1e3), flush=True)
comp = lib.alloc(K * 48)
ms = timed(lambda: lib.check(lib.c.cpg_g1_compress(jac.ptr, K, comp.ptr)))
out["compress_per_s"] = K / ms * 1e3
print("compress: %.2f ms -> %.3g /s" % (ms, K / ms * 1e3), flush=True)
affo = lib.alloc(K * rt.AFF); err = lib.alloc(K)
ms = timed(lambda: lib.check(lib.c.cpg_g1_decompress(comp.ptr, K, 0, affo.ptr, err.ptr)))
out["decompress_per_s"] = K / ms * 1e3
print("decompress: %.2f ms -> %.3g /s" % (ms, K / ms * 1e3), flush=True)
ms = timed(lambda: lib.add(jac, jac, K))
print("add (doubling branch): %.2f ms -> %.3g /s" % (ms, K / ms * 1e3), flush=True)

for B, n, cs in ((1024, 128, (4, 5, 6)), (128, 627, (6, 7, 8)), (1, 1 << 14, (9, 11)),):
    if B * n > K:
        continue
    sc = rand_scalars(B * n)
    for c in cs:
        res = lib.alloc(B * rt.JAC)
        ms = timed(lambda: lib.msm_batched(aff, n, sc, B, n, c, out=res))
        print("msm B=%d n=%d c=%d: %.2f ms -> %.3g Mpoints/s" % (B, n, c, ms, B * n / ms / 1e3), flush=True)
        out["msm_B%d_n%d_c%d_ms" % (B, n, c)] = ms
for c in (8, 12):
    table = lib.fixed_table(aff, 131, c)
    B = 1024
    sc = rand_scalars(B * 131)
    res = lib.alloc(B * rt.JAC)
    ms = timed(lambda: lib.msm_fixed_batched(table, sc, B, out=res))
    print("fixed msm nb=131 B=%d c=%d (%d MB table): %.2f ms -> %.3g Mpoints/s" % (B, c, table.nbytes >> 20, ms, B * 131 / ms / 1e3), flush=True)
    out["fixed_c%d_ms" % c] = ms
    table.free()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)

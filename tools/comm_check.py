"""Multi-GPU check of the in-library communicator (include/cpg.h: cpg_comm_*, cpg_g1_msm_sharded), one process per GPU:

    python tools/comm_check.py [--gpus 2] [--terms 65536]

Spawns the ranks itself (no torchrun, no torch): every rank builds the same seeded MSM instance on its own GPU, runs
cpg_g1_msm_sharded (window slice -> ncclAllGather on device buffers -> Horner) and compares the compressed result with
its own single-GPU MSM and, on rank 0, with the oracle; then an allgather_bytes / max_over_ranks / gather_verdicts round.
Prints one JSON line per rank 0 with the timings."""
import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, n, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), CPG_DEVICE=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from curdleproofs_pie_b200 import comm, msm, runtime as rt, sharding

    lib = rt.get_lib()
    comm.init_from_env(lib)
    assert lib.c.cpg_comm_world() == world and lib.c.cpg_comm_rank() == rank
    rng = random.Random(5)
    nuniq = 64
    gen = lib.generator()
    gens = lib.alloc(nuniq * rt.JAC)
    for i in range(nuniq):
        lib.check(lib.c.cpg_d2d(gens.ptr + i * rt.JAC, gen.ptr, rt.JAC))
    uniq = [rng.randrange(1, rt.R_ORDER) for _ in range(nuniq)]
    uaff = lib.jac_to_aff(lib.mul(gens, lib.upload(rt.scalars_to_bytes(uniq)), nuniq), nuniq)
    bases = lib.alloc(n * rt.AFF)
    for i in range(0, n, nuniq):
        lib.check(lib.c.cpg_d2d(bases.ptr + i * rt.AFF, uaff.ptr, min(nuniq, n - i) * rt.AFF))
    ks = [rng.randrange(rt.R_ORDER) for _ in range(n)]
    scalars = lib.upload(rt.scalars_to_bytes(ks))
    c = int(lib.c.cpg_msm_pick_window(n))
    single = lib.alloc(rt.JAC)
    lib.check(lib.c.cpg_g1_msm_batched(bases.ptr, 0, scalars.ptr, 1, n, c, single.ptr))
    out = msm.msm_large(lib, bases, scalars, n, window=c)
    assert lib.compress_jac(out, 1) == lib.compress_jac(single, 1), "sharded MSM differs from the single-GPU MSM"
    if rank == 0:
        from oracle import cref_binding

        cref = cref_binding.load()
        agg = [0] * nuniq
        for i, k in enumerate(ks):
            agg[i % nuniq] = (agg[i % nuniq] + k) % rt.R_ORDER
        enc = lib.compress_aff(uaff, nuniq)
        blobs = [cref.decompress(enc[48 * i:48 * i + 48], False) for i in range(nuniq)]
        assert lib.compress_jac(out, 1) == cref.compress(cref.msm(blobs, agg)), "sharded MSM differs from the oracle"
    times = {}
    for name, fn in (("single", lambda: lib.c.cpg_g1_msm_batched(bases.ptr, 0, scalars.ptr, 1, n, c, single.ptr)),
                     ("sharded", lambda: lib.c.cpg_g1_msm_sharded(bases.ptr, scalars.ptr, n, c, out.ptr))):
        for _ in range(3):
            lib.check(fn())
        lib.sync()
        lib.timer_start()
        for _ in range(10):
            lib.check(fn())
        times[name] = lib.timer_stop() / 10
    parts = comm.allgather_bytes(lib, bytes([rank]) * 5)
    assert parts == [bytes([r]) * 5 for r in range(world)]
    assert sharding.max_over_ranks(10.0 + rank, lib=lib) == 10.0 + world - 1
    total = 11
    lo, hi = sharding.shard_range(total, rank, world)
    full = sharding.gather_verdicts(bytes((i * 7 + 3) % 2 for i in range(lo, hi)), total, rank, world, lib=lib)
    assert full == bytes((i * 7 + 3) % 2 for i in range(total))
    # ONE proof over all ranks (BASELINE config 5's decomposition): bytes equal the reference's golden proof, verdicts too
    import prove_cases as prc
    import verify_cases as vc
    import large_cases as lc

    prc.check_prove(lib, "shuffle_N128_seed4096.json", copies=2, sharded=True)
    prc.check_prove(lib, "shuffle_N16_seed77.json", sharded=True, transcript="device")
    vc.check_batch(lib, "shuffle_N128_seed4096.json", sharded=True, transcript_on_device=False)
    vc.check_batch(lib, "shuffle_N16_seed77.json", sharded=True, transcript_on_device=True)
    big = lc.check_large(lib, "large_N1024_seed6024.json", fixed_window=8, sharded=True)
    times["large_N1024_prove"] = big["prove_s"] * 1e3
    slow = {k: sharding.max_over_ranks(v, lib=lib) for k, v in times.items()}
    lib.c.cpg_comm_free()
    q.put((rank, slow, int(lib.c.cpg_comm_nccl_version())))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--terms", type=int, default=1 << 16)
    ap.add_argument("--port", type=int, default=29611)
    a = ap.parse_args()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    t0 = time.time()
    procs = [ctx.Process(target=worker, args=(r, a.gpus, a.port, a.terms, q)) for r in range(a.gpus)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    print(json.dumps({"check": "cpg_g1_msm_sharded == single-GPU == oracle; allgather / max / verdict gather through libcpg's NCCL communicator",
                      "gpus": a.gpus, "n": a.terms, "ms_single_gpu": res[0][1]["single"], "ms_sharded": res[0][1]["sharded"], "sharded_proofs": "N128 / N16 golden bytes + verdicts, N1024 digests: equal the reference's", "ms_large_N1024_prove": res[0][1]["large_N1024_prove"],
                      "nccl_version": res[0][2], "wall_s": time.time() - t0}))


if __name__ == "__main__":
    main()

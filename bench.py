#!/usr/bin/env python
"""bench.py - the hot path's headline measurement (one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload msm]

A "step" is one pass of the hot path over one batch of synthetic input:
  workload verify (default): B Whisk-size (n = 128, ell = 124) shuffle proofs verified per GPU per
                 step (BASELINE config 4's per-GPU share).  The proofs are the reference-generated
                 golden proof (tests/golden, made by the UNMODIFIED reference) replicated B times
                 with lane-unique batching weights and 1 lane in 64 corrupted; every step's verdict
                 bitmap is checked against the expected one.  metric = verifications/s.
                   value : device side on HBM-resident inputs (decompress, D/A', per-proof MSM,
                           verdict) timed with CUDA events - cpg_verify_replay_device
                   e2e   : cpg_verify_batch with HOST buffers in and verdict bytes out, wall clock
                           (H2D, host transcript + Fr algebra on all host threads, D2H inside)
  workload msm : B independent G1 MSMs of n = 128 terms each (BASELINE config 2/3 shape: the
                 Whisk-size MSM behind every sub-proof), per-MSM bases, uniform scalars < r.
Sharding (--gpus N under torchrun): each rank owns B MSMs on its own GPU, no data-path collective
 -> "scaling": "weak".  value = units of all ranks / max-over-ranks device time.

  value        inputs resident in HBM (affine bases + scalars), output Jacobian points in HBM
  e2e          through the C ABI with HOST buffers: pinned compressed bases (48 B) + scalars (32 B)
               -> H2D -> decompress -> MSM -> compress -> D2H of 48 B per MSM, all inside the timed region
  roofline     dominant kernel (BucketAccumulate) against the integer (IMAD) pipe: algorithmic
               32x32->64 MACs / its device time (CUDA events on the launching stream, live), peak
               = saturating IMAD.WIDE.U32 microbenchmark measured in the same run
  cpu_baseline the oracle's C restatement (oracle/cref: Pippenger with arkworks' window rule), one
               host core, bounded sample
--impl reference: the same workload on the host cores (all of them) through the oracle port - the
  reference's arithmetic lives in the py_arkworks_bls12381 wheel, which is not installable offline.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dropin")]

R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# DRAM bytes per unit from the committed `ncu --set full` capture (profiles/r01_ncu_full_v10_key_metrics.txt:
# dram__bytes_read.sum + dram__bytes_write.sum of one launch over 4096 proofs), scaled to the launch the bench times
NCU_DRAM_BYTES = {"Decompress": (120.727e6 + 329.198e6) / (4096 * 586),           # per point (algorithmic: 48 in + 96 + 1 out)
                  "BucketAccumulate": (654.527e6 + 1826.430e6) / (4096 * 586 * 37)}  # per (term, window) at c = 7, one MSM per proof
MAC_PER_MODMUL = 300     # 12-limb Montgomery product: 2*12^2 + 12
SQR_MAC = 222            # dedicated Montgomery squaring: 78 product + 144 reduction MACs


def msm_model(n, c):
    """Algorithmic Fq products of one n-term MSM with window c (DESIGN.md 'work model')."""
    W = (256 + c - 1) // c
    NB = 1 << (c - 1)
    return {"W": W, "NB": NB, "bucket_accumulate": n * W * 10, "window_reduce": W * NB * 2 * 14,
            "horner": (W - 1) * (c * 9 + 14)}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def rand_scalar_bytes(rng, k):
    return b"".join(rng.randrange(R_ORDER).to_bytes(32, "little") for _ in range(k))


def device_random_points(lib, rt, rng, k):
    """k random G1 points s_i * G computed on the device; returns a Jacobian DevBuf."""
    gen = lib.generator()
    gens = lib.alloc(k * rt.JAC)
    lib.check(lib.c.cpg_d2d(gens.ptr, gen.ptr, rt.JAC))
    have = 1
    while have < k:
        cnt = min(have, k - have)
        lib.check(lib.c.cpg_d2d(gens.ptr + have * rt.JAC, gens.ptr, cnt * rt.JAC))
        have += cnt
    ks = lib.upload(rand_scalar_bytes(rng, k))
    return lib.mul(gens, ks, k)


# ------------------------------------------------------------------------------------------------
def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "shuffle_N128_seed4096.json")) as f:
        return json.load(f)


def pick_window(n):
    best, bc = None, 4
    for c in range(3, 17):
        m = msm_model(n, c)
        cost = m["bucket_accumulate"] + m["window_reduce"] + m["horner"]
        if best is None or cost < best:
            best, bc = cost, c
    return bc


def make_verify_batch(case, B, every=64):
    """B lanes of (inputs, proof, expected verdict): the golden proof, 1 lane in `every` corrupted."""
    cat = lambda k: b"".join(bytes.fromhex(h) for h in case[k])  # noqa: E731
    R, S, T, U = cat("vec_R"), cat("vec_S"), cat("vec_T"), cat("vec_U")
    good_in = R + S + T + U
    proof = bytes.fromhex(case["M"]) + bytes.fromhex(case["proof"])
    bad_in = S + R + T + U                                        # swapped R/S (cp/test_curdleproofs.py:643-650)
    bad_proof = bytearray(proof); bad_proof[48 * 9 + 7] ^= 1; bad_proof = bytes(bad_proof)   # flipped bit in cm_U / B
    ins, prs, exp = [], [], bytearray(B)
    for i in range(B):
        if every and i % every == every - 1:
            if (i // every) % 2 == 0:
                ins.append(bad_in); prs.append(proof)
            else:
                ins.append(good_in); prs.append(bad_proof)
        else:
            ins.append(good_in); prs.append(proof); exp[i] = 1
    return b"".join(ins), b"".join(prs), bytes(exp)


def run_ours_verify(args, rank, world, dist):
    import ctypes

    from curdleproofs_pie_b200 import runtime as rt
    from curdleproofs_pie_b200 import whisk

    lib = rt.get_lib()
    assert lib.backend == "cuda-sm_100a", "bench must run on the CUDA library, got " + lib.backend
    case = load_golden()
    B, ell, n = args.batch, 124, 128
    NV, NF = 4 * ell + 1 + 18 + 10 * 7 + 1, n + 3
    c = args.window or pick_window(NV)
    # host threads stage the wire bytes (and run the transcript with --transcript host): share the cores between the ranks
    host_threads = args.host_threads or max(1, (os.cpu_count() or 1) // world)
    ver = whisk.BatchVerifier(bytes.fromhex(case["crs"]), ell, fixed_window=args.fixed_window, host_threads=host_threads)
    ver.set_window(c)
    ver.set_transcript(args.transcript == "device")
    ver.set_streams(args.streams)
    ver.set_group(args.group, args.group_window)
    inputs, proofs, expected = make_verify_batch(case, B, args.corrupt_every)
    out = ctypes.create_string_buffer(B)

    def barrier():
        lib.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(ms):
        from curdleproofs_pie_b200 import sharding

        return sharding.max_over_ranks(ms, dist, device="cuda")

    def step_e2e():
        lib.check(lib.c.cpg_verify_batch(ver.handle, inputs, proofs, B, out), "cpg_verify_batch")
        assert out.raw[:B] == expected, "verdict bitmap differs from the expected one"

    def step_dev():
        lib.check(lib.c.cpg_verify_replay_device(ver.handle, None), "cpg_verify_replay_device")

    warm = max(args.warmup, 3)
    step_e2e()
    for _ in range(warm):
        step_dev()
    barrier()
    sampler = ClockSampler(lib.device) if rank == 0 else None
    launches0 = lib.launch_count()
    barrier()
    lib.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms = lib.timer_stop()
    barrier()
    launches = lib.launch_count() - launches0
    # per-kernel device times for the roofline: a separate pass on ONE stream with an event pair around
    # every launch (with several sub-batch streams the kernels overlap and per-kernel times are meaningless)
    ver.set_streams(1)
    step_dev()
    lib.sync()
    lib.profile(True)
    for _ in range(args.steps):
        step_dev()
    prof = lib.profile_report()
    lib.profile(False)
    ver.set_streams(args.streams)
    # verdicts of the replayed device pass (no host-side rejects in a replay: all lanes decode)
    lib.check(lib.c.cpg_verify_replay_device(ver.handle, out))
    assert out.raw[:B] == expected, "device replay verdicts differ"
    ms = max_over_ranks(ms)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    lib.sync()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None

    peak_mac, _ = lib.bench_int_pipe(0, 20000)
    fq_rate, _ = lib.bench_int_pipe(2, 2000)
    ver_rechecked, ver_group = ver.rechecked(), ver.group()
    # ---- algorithmic Fq products per step, kernel by kernel (DESIGN.md "work model") ----
    # decompress: x^3 + 4, the 376-squaring / 81-product square-root chain, y^2 check, Montgomery in/out
    DEC_MAC = 378 * SQR_MAC + 86 * MAC_PER_MODMUL
    model = msm_model(NV, c)
    per_proof = B if ver_group <= 1 else ver_rechecked                 # proofs on their own MSM
    sub = B                                                            # the per-kernel pass runs on ONE stream: one sub-batch
    if ver_group > 1:
        cg = args.group_window or int(lib.c.cpg_msm_pick_window_batched(max(1, sub // ver_group), ver_group * NV))
        gmodel = msm_model(ver_group * NV, cg)
        ngroups = B // ver_group
    else:
        cg, gmodel, ngroups = None, {"bucket_accumulate": 0, "window_reduce": 0, "horner": 0}, 0
    alg_mac = {
        "Decompress": B * (NV - 1) * DEC_MAC,
        "BucketAccumulate": (per_proof * model["bucket_accumulate"] + ngroups * gmodel["bucket_accumulate"]) * MAC_PER_MODMUL,
        "WindowReduce": per_proof * model["window_reduce"] * MAC_PER_MODMUL,
        "ReduceLevel": ngroups * gmodel["window_reduce"] * MAC_PER_MODMUL,
        "FixedMsmWindow": (per_proof + ngroups) * NF * 22 * 10 * MAC_PER_MODMUL + B * 2 * 32 * 10 * MAC_PER_MODMUL,
    }
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        ms_step = v["ms"] / args.steps
        ent = {"ms_per_step": ms_step, "launches_per_step": v["launches"] / args.steps}
        if k in alg_mac and ms_step > 0:
            ent["achieved_gmac_s"] = alg_mac[k] / (ms_step * 1e-3) / 1e9
            ent["frac"] = alg_mac[k] / (ms_step * 1e-3) / peak_mac if peak_mac else None
        kernels[k] = ent
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
    dom_ms = prof[dom]["ms"] / args.steps
    dom_launch_ms = prof[dom]["ms"] / max(1, prof[dom]["launches"])
    dom_mac_launch = alg_mac.get(dom, 0) * args.steps / max(1, prof[dom]["launches"])
    achieved = dom_mac_launch / (dom_launch_ms * 1e-3) if dom_launch_ms else 0.0
    # whole device pass in algorithmic MACs: the kernels above + Horner + D (2 scalar-muls per proof)
    mac_step = sum(alg_mac.values()) + (per_proof * model["horner"] + ngroups * gmodel["horner"] + B * 2 * 2900) * MAC_PER_MODMUL
    roofline = {
        "bound": "int_pipe", "kernel": dom, "achieved": achieved / 1e9, "peak": peak_mac / 1e9, "unit": "GMAC/s",
        "frac": achieved / peak_mac if peak_mac else None,
        "traffic": (NCU_DRAM_BYTES["Decompress"] * B * (NV - 1) * args.steps / max(1, prof[dom]["launches"])) if dom == "Decompress" else None,
        "traffic_note": "bytes per launch = the committed ncu capture's DRAM bytes per point (187 B read + written; 145 B algorithmic) x the points of one launch: ~13 GB/s against 6.5 TB/s, HBM is idle",
        "peak_source": "data-dependent IMAD.WIDE.U32 chains measured in this run (32 lanes/clk/SM; MEASURED_PEAKS.json has no integer-pipe entry)",
        "kernel_ms_per_launch": dom_launch_ms, "kernel_share_of_step": dom_ms * args.steps / total_kernel_ms if total_kernel_ms else None,
        "algorithmic_mac_per_launch": dom_mac_launch, "mac_per_modmul": MAC_PER_MODMUL, "mac_per_modsqr": SQR_MAC,
        "fq_mul_chain_per_s": fq_rate, "whole_step_mac_per_s": mac_step / (ms / args.steps * 1e-3),
        "whole_step_frac_of_peak": mac_step / (ms / args.steps * 1e-3) / peak_mac if peak_mac else None,
        "group_window": cg, "kernels": kernels,
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
    }
    cpu = cpu_verify_rate(case, sample=args.cpu_sample_verify, procs=1)
    prove_side = None
    if args.prove_batch and world == 1:
        ver.close()
        pr = measure_prove(args, lib, case, args.prove_batch, 2)
        prove_side = {"metric": "curdleproofs_prove_per_s_n128_batched", "value": pr["B"] * pr["steps"] / (pr["ms"] * 1e-3),
                      "e2e": pr["B"] * pr["steps"] / (pr["ms_e2e"] * 1e-3), "unit": "proofs/s", "B": pr["B"], "ms_per_step": pr["ms"] / pr["steps"],
                      "note": "side measurement on this GPU (python bench.py --workload prove gives the full line)"}
    total = B * world
    return {
        "prove": prove_side,
        "metric": "curdleproofs_verify_per_s_n128_batched", "value": total * args.steps / (ms * 1e-3), "unit": "verifications/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq, 255-bit Fr)", "data": "synthetic",
        "config": {"workload": "batch verification of Whisk-size (n=128, ell=124) curdleproofs, B=%d proofs per GPU per step; "
                               "reference-generated golden proof replicated, lane-unique weights, %s lanes corrupted, verdicts checked" % (B, "1/%d" % args.corrupt_every if args.corrupt_every else "no"),
                   "group": args.group, "group_in_use": ver_group, "rechecked_per_step": ver_rechecked,
                   "B_per_gpu": B, "n": n, "var_terms": NV, "fixed_terms": NF, "window": c, "transcript": args.transcript, "streams": args.streams,
                   "l2": "inputs_larger_than_l2 (%.0f MB wire + scalars per step)" % (B * (NV * 80 + NF * 32) / 1e6), "sharding": "per-proof, no collective"},
        "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "verifications/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * (NV * 48 + 7 * 32 + (0 if args.transcript == "device" else 64 + NV * 32 + NF * 32 + 1)),
                "d2h_bytes_per_step": B * (1 + (0 if args.transcript == "device" else 96 + NV + 1)),
                "host_threads": host_threads},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }


def _cpu_verify_worker(job):
    case, count = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import ark_surface, cref_binding, merlin_py
    from oracle.shuffle_ref import ShuffleRef

    ark_surface.set_backend("c")
    merlin_py.use_native_keccak(cref_binding.load().keccak_f1600)
    G1Point, Scalar = ark_surface.G1Point, ark_surface.Scalar
    ctx = ShuffleRef(G1Point, Scalar)
    ell = case["N"] - 4
    crs = ctx.crs_from_bytes(bytes.fromhex(case["crs"]), ell)
    proof = bytes.fromhex(case["proof"])
    t0 = time.perf_counter()
    for _ in range(count):
        # the Whisk boundary decodes every tracker per call (whisk_interface.py:96-100)
        dec = lambda lst: [G1Point.from_compressed_bytes_unchecked(bytes.fromhex(h)) for h in lst]  # noqa: E731
        R_, S_, T_, U_ = dec(case["vec_R"]), dec(case["vec_S"]), dec(case["vec_T"]), dec(case["vec_U"])
        M = G1Point.from_compressed_bytes_unchecked(bytes.fromhex(case["M"]))
        assert ctx.is_valid(crs, R_, S_, T_, U_, M, proof)
    return time.perf_counter() - t0


def cpu_verify_rate(case, sample, procs):
    """The oracle's per-proof restatement of the reference verifier (oracle/shuffle_ref.py on the C
    arithmetic of oracle/cref, double-and-add scalar-muls like arkworks' G1Projective * Fr, C Keccak
    under the Python Merlin framing) on `procs` host cores."""
    from oracle import cref_binding

    cref_binding.build()
    per = max(1, sample // procs)
    if procs == 1:
        secs = _cpu_verify_worker((case, per))
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(procs) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_verify_worker, [(case, per)] * procs)
            secs = time.perf_counter() - t0
    done = per * procs
    return {"value": done / secs, "unit": "verifications/s", "cores": procs, "kind": "port",
            "sample": "%d verifications of the golden n=128 proof (oracle/shuffle_ref.py, C arithmetic)" % done, "seconds": secs}


def run_reference_verify(args, rank, world):
    if rank != 0:
        return None
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    case = load_golden()
    vals, secs = [], 0.0
    for i in range(args.warmup + args.steps):
        r = cpu_verify_rate(case, sample=cores * 2, procs=cores)
        if i >= args.warmup:
            vals.append(r["value"]); secs += r["seconds"]
    value = sum(vals) / len(vals)
    return {
        "impl": "reference", "metric": "curdleproofs_verify_per_s_n128_batched", "value": value, "unit": "verifications/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": {"workload": "verification of Whisk-size (n=128) curdleproofs; each step = %d verifications spread over %d host cores" % (cores * 2, cores), "n": 128},
        "cpu_baseline": {"value": value, "unit": "verifications/s", "cores": cores, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": value, "unit": "verifications/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "py_arkworks_bls12381 (the reference's Rust arithmetic) is not installable offline; this arm times the oracle port of the reference verifier",
    }



# ------------------------------------------------------------------------------------------------
def make_prove_batch(case, prover, B, seed):
    import random as pyrandom

    rng = pyrandom.Random(seed)
    ell = case["N"] - 4
    pre = b"".join(bytes.fromhex(h) for h in case["vec_R"] + case["vec_S"])
    perms, ks, rands = [], bytearray(), bytearray()
    for _ in range(B):
        p = list(range(ell)); rng.shuffle(p)
        perms.extend(p)
        ks += (rng.getrandbits(248) + 1).to_bytes(32, "little")                 # < 2^249 < r, non-zero
        rands += b"".join((rng.getrandbits(248) + 1).to_bytes(32, "little") for _ in range(prover.n_rand))
    return pre * B, perms, bytes(ks), bytes(rands)


def measure_prove(args, lib, case, B, steps, dist=None):
    """Returns a dict with device (resident inputs, CUDA events) and e2e (host buffers, wall clock)
    proofs/s for B proofs per step; the produced proofs are checked by the batched verifier."""
    import array
    import ctypes

    from curdleproofs_pie_b200 import whisk

    ell = case["N"] - 4
    crs = bytes.fromhex(case["crs"])
    prover = whisk.BatchProver(crs, ell, fixed_window=args.fixed_window)
    c = args.prove_window or pick_window(ell)
    prover.set_window(c)
    prover.set_lanes(args.prove_lanes)
    prover.set_table_window(args.table_window)
    inputs, perms, ks, rands = make_prove_batch(case, prover, B, 4242)
    perm_arr = array.array("I", perms)
    pbuf = (ctypes.c_uint32 * len(perm_arr)).from_buffer(perm_arr)
    out_tu = ctypes.create_string_buffer(B * 2 * ell * 48)
    out_pr = ctypes.create_string_buffer(B * prover.proof_len)
    status = ctypes.create_string_buffer(B)

    def step_e2e():
        lib.check(lib.c.cpg_prove_batch(prover.handle, inputs, pbuf, ks, rands, B, out_tu, out_pr, status), "cpg_prove_batch")

    def step_dev():
        lib.check(lib.c.cpg_prove_replay_device(prover.handle), "cpg_prove_replay_device")

    step_e2e()
    assert not any(status.raw[:B])
    # every produced proof must be accepted by the batched verifier
    ver = whisk.BatchVerifier(crs, ell)
    w = 2 * ell * 48
    pre = inputs[:w]
    tus, prs = out_tu.raw, out_pr.raw
    nchk = min(B, 256)
    verdicts = ver.verify([pre + tus[i * w:(i + 1) * w] for i in range(nchk)], [prs[i * prover.proof_len:(i + 1) * prover.proof_len] for i in range(nchk)])
    assert all(verdicts), "a generated proof was rejected"
    ver.close()
    for _ in range(2):
        step_dev()
    lib.sync()
    launches0 = lib.launch_count()
    lib.timer_start()
    for _ in range(steps):
        step_dev()
    ms = lib.timer_stop()
    launches = lib.launch_count() - launches0
    # per-kernel times: a separate pass with ONE lane (with several lanes the kernels overlap on streams)
    prover.set_lanes(1)
    step_e2e()
    lib.sync()
    lib.profile(True)
    for _ in range(steps):
        step_dev()
    prof = lib.profile_report()
    lib.profile(False)
    prover.set_lanes(args.prove_lanes)
    step_e2e()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    lib.sync()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    prover.close()
    return {"B": B, "window": c, "ms": ms, "ms_e2e": ms_e2e, "steps": steps, "launches": launches, "prof": prof,
            "h2d": B * (2 * ell * 48 + ell * 4 + 32 + prover.n_rand * 32), "d2h": B * (2 * ell * 48 + prover.proof_len + (21 + 70) * 48 + 2 * ell)}


def _cpu_prove_worker(job):
    case, count = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import random as pyrandom

    from oracle import ark_surface, cref_binding, merlin_py
    from oracle.shuffle_ref import ShuffleRef

    ark_surface.set_backend("c")
    merlin_py.use_native_keccak(cref_binding.load().keccak_f1600)
    G1Point, Scalar = ark_surface.G1Point, ark_surface.Scalar
    ctx = ShuffleRef(G1Point, Scalar, rng=pyrandom.Random(99))
    ell = case["N"] - 4
    crs = ctx.crs_from_bytes(bytes.fromhex(case["crs"]), ell)
    dec = lambda lst: [G1Point.from_compressed_bytes_unchecked(bytes.fromhex(h)) for h in lst]  # noqa: E731
    t0 = time.perf_counter()
    for _ in range(count):
        vec_R, vec_S = dec(case["vec_R"]), dec(case["vec_S"])             # the Whisk boundary decodes per call
        perm = list(range(ell)); ctx.rng.shuffle(perm)
        k = ctx.rand()
        vec_T, vec_U, M, m_bl = ctx.shuffle_and_commit(crs, vec_R, vec_S, perm, k)
        ctx.prove(crs, vec_R, vec_S, vec_T, vec_U, M, perm, k, m_bl)
    return time.perf_counter() - t0


def cpu_prove_rate(case, sample, procs):
    from oracle import cref_binding

    cref_binding.build()
    per = max(1, sample // procs)
    if procs == 1:
        secs = _cpu_prove_worker((case, per))
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(procs) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_prove_worker, [(case, per)] * procs)
            secs = time.perf_counter() - t0
    done = per * procs
    return {"value": done / secs, "unit": "proofs/s", "cores": procs, "kind": "port",
            "sample": "%d n=128 proofs (oracle/shuffle_ref.py incl. the reference's self-check MSMs, C arithmetic)" % done, "seconds": secs}


def run_ours_prove(args, rank, world, dist):
    from curdleproofs_pie_b200 import runtime as rt
    from curdleproofs_pie_b200 import sharding

    lib = rt.get_lib()
    assert lib.backend == "cuda-sm_100a"
    case = load_golden()
    B = args.batch if args.batch != 8192 else 4096          # BASELINE config 3: 4096 proofs per batch
    sampler = ClockSampler(lib.device) if rank == 0 else None
    if dist is not None:
        dist.barrier()
    r = measure_prove(args, lib, case, B, args.steps)
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = sharding.max_over_ranks(r["ms"], dist, device="cuda")
    ms_e2e = sharding.max_over_ranks(r["ms_e2e"], dist, device="cuda")
    if rank != 0:
        return None
    peak_mac, _ = lib.bench_int_pipe(0, 20000)
    prof = r["prof"]
    # Algorithmic MACs per step.  Variable-base MSMs: R', S', B_t, B_u over all ell leaves and, per SameMSM
    # round, L_T L_U R_T R_U over half of them (a fold only rescales leaf weights): 4 ell + 4 lg ell/2 non-zero
    # terms per proof in 4 + 4 lg MSM instances.  The fixed-base kernel skips structurally zero coefficients,
    # whose count is not modelled here, so it carries no roofline entry and the whole-step figure is a lower bound.
    ell_, lg_ = 124, 7
    var = msm_model(1, r["window"])
    tw = args.table_window
    tu_terms, tu_inst = 2 * ell_ + 4 * lg_ * (ell_ // 2), 2 + 4 * lg_          # B_t, B_u + L/R of T and U per SameMSM round
    rs_terms, rs_inst = 2 * ell_, 2                                          # R', S'
    bucket_terms, bucket_inst = (rs_terms, rs_inst) if tw else (rs_terms + tu_terms, rs_inst + tu_inst)
    DEC_MAC = 378 * SQR_MAC + 86 * MAC_PER_MODMUL
    INV_MAC = 377 * SQR_MAC + 83 * MAC_PER_MODMUL
    alg_mac = {
        "BucketAccumulate": B * bucket_terms * var["W"] * 10 * MAC_PER_MODMUL,
        "WindowReduce": B * bucket_inst * var["window_reduce"] * MAC_PER_MODMUL,
        # GLV shuffle: 128 doublings (2M+5S) + ~62 Jacobian additions (11M+5S) + table + phi, then one inversion to affine
        "ProveShuffle": B * 2 * ell_ * (132 * (2 * MAC_PER_MODMUL + 5 * SQR_MAC) + 70 * (11 * MAC_PER_MODMUL + 5 * SQR_MAC) + 40 * MAC_PER_MODMUL + INV_MAC),
        "Decompress": B * 2 * ell_ * DEC_MAC,
    }
    if tw:
        Wt, TS = (256 + tw - 1) // tw, 1 << (tw - 1)
        alg_mac["VarTableMsmWindow"] = B * tu_terms * Wt * 10 * MAC_PER_MODMUL
        # per entry: mixed Jacobian addition (7M+4S), prefix product, 2 back-substitution products, 1S+3M to affine; one inversion per base
        alg_mac["VarTableBuild"] = B * 2 * ell_ * (TS * (13 * MAC_PER_MODMUL + 5 * SQR_MAC) + INV_MAC)
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        ms_step = v["ms"] / args.steps
        ent = {"ms_per_step": ms_step, "launches_per_step": v["launches"] / args.steps}
        if k in alg_mac and ms_step > 0:
            ent["achieved_gmac_s"] = alg_mac[k] / (ms_step * 1e-3) / 1e9
            ent["frac"] = alg_mac[k] / (ms_step * 1e-3) / peak_mac if peak_mac else None
        kernels[k] = ent
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    modelled = [k for k in prof if k in alg_mac]
    dom = max(modelled, key=lambda k: prof[k]["ms"]) if modelled else "BucketAccumulate"
    ba = prof.get(dom, {"ms": 0.0, "launches": 1})
    ba_launch_ms = ba["ms"] / max(1, ba["launches"])
    ba_mac_launch = alg_mac.get(dom, 0) * args.steps / max(1, ba["launches"])
    cpu = cpu_prove_rate(case, sample=args.cpu_sample_prove, procs=1)
    total = B * world
    return {
        "metric": "curdleproofs_prove_per_s_n128_batched", "value": total * args.steps / (ms * 1e-3), "unit": "proofs/s",
        "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq, 255-bit Fr)", "data": "synthetic",
        "config": {"workload": "batch generation of Whisk-size (n=128, ell=124) curdleproofs, B=%d proofs per GPU per step in lock-step; "
                               "pre-shuffle trackers from the reference-generated fixture, per-lane permutation / k / blinders; 256 of the proofs re-checked by the batched verifier" % B,
                   "B_per_gpu": B, "n": 128, "window_var": r["window"], "window_fixed": args.fixed_window or 12, "lanes": args.prove_lanes, "sharding": "per-proof, no collective",
                   "l2": "inputs_larger_than_l2 (%.0f MB per step)" % (r["h2d"] / 1e6)},
        "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
        "gpu_launches": r["launches"], "clocks": clocks,
        "roofline": {"bound": "int_pipe", "kernel": dom, "achieved": ba_mac_launch / (ba_launch_ms * 1e-3) / 1e9 if ba_launch_ms else 0.0,
                     "peak": peak_mac / 1e9, "unit": "GMAC/s", "frac": (ba_mac_launch / (ba_launch_ms * 1e-3) / peak_mac) if ba_launch_ms and peak_mac else None,
                     "traffic": None, "kernel_ms_per_launch": ba_launch_ms, "algorithmic_mac_per_launch": ba_mac_launch,
                     "kernel_share_of_step": ba["ms"] / total_kernel_ms if total_kernel_ms else None,
                     "whole_step_frac_of_peak_lower_bound": sum(alg_mac.values()) / (ms / args.steps * 1e-3) / peak_mac if peak_mac else None,
                     "peak_source": "data-dependent IMAD.WIDE.U32 chains measured in this run",
                     "kernels": kernels,
                     "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}},
        "cpu_baseline": cpu,
    }



# ------------------------------------------------------------------------------------------------
def run_ours_msm_large(args, rank, world, dist):
    """ONE n-term MSM (G1Point.multiexp_unchecked, BASELINE config 2's sweep / config 5's building block).
    N = 1: all windows on one GPU.  N > 1: Pippenger windows split across ranks (every rank holds all
    bases and scalars), one NCCL all-gather of the W window sums (W*144 B), Horner on every rank:
    strong scaling.  The result is compared with the oracle on a structured instance."""
    from curdleproofs_pie_b200 import msm as msm_mod
    from curdleproofs_pie_b200 import runtime as rt
    from curdleproofs_pie_b200 import sharding

    lib = rt.get_lib()
    assert lib.backend == "cuda-sm_100a"
    n = args.n if args.n != 128 else 1 << 20
    rng = random.Random(7)                                  # identical inputs on every rank
    nuniq = 4096
    base_jac = device_random_points(lib, rt, rng, nuniq)
    uaff = lib.jac_to_aff(base_jac, nuniq)
    bases = lib.alloc(n * rt.AFF)                           # n bases drawn from 4096 distinct points
    have = 0
    while have < n:
        cnt = min(nuniq, n - have)
        lib.check(lib.c.cpg_d2d(bases.ptr + have * rt.AFF, uaff.ptr, cnt * rt.AFF))
        have += cnt
    sc_bytes = bytearray()                                  # seeded: every rank must hold the same scalars
    while len(sc_bytes) < 32 * n:                           # (randbytes is limited to 2^28 bytes per call)
        sc_bytes += rng.randbytes(min(1 << 24, 32 * n - len(sc_bytes)))
    sc_bytes[31::32] = bytes(b & 0x3F for b in sc_bytes[31::32])      # < 2^254 < r
    scalars = lib.upload(bytes(sc_bytes))
    c = args.window or int(lib.c.cpg_msm_pick_window(n))
    dev = "cuda" if dist is not None else "cpu"

    def step():
        return msm_mod.msm_large(lib, bases, scalars, n, window=c, dist=dist, device=dev)

    def barrier():
        lib.sync()
        if dist is not None:
            dist.barrier()

    out = step()
    # parity on this very instance: aggregate the scalars per distinct base on the host, oracle MSM of 4096 terms
    if rank == 0:
        from oracle import cref_binding

        cref = cref_binding.load()
        agg = [0] * nuniq
        for i in range(n):
            agg[i % nuniq] += int.from_bytes(sc_bytes[32 * i:32 * i + 32], "little")
        enc = lib.compress_aff(uaff, nuniq)
        blobs = [cref.decompress(enc[48 * i:48 * i + 48], False) for i in range(nuniq)]
        want = cref.compress(cref.msm(blobs, [a % R_ORDER for a in agg]))
        assert lib.compress_jac(out, 1) == want, "large MSM differs from the oracle"
    for _ in range(max(args.warmup, 3) - 1):
        step()
    barrier()
    sampler = ClockSampler(lib.device) if rank == 0 else None
    lib.profile(True)
    launches0 = lib.launch_count()
    barrier()
    t0 = time.perf_counter()
    lib.timer_start()
    for _ in range(args.steps):
        step()
    ms_dev = lib.timer_stop()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    launches = lib.launch_count() - launches0
    prof = lib.profile_report()
    lib.profile(False)
    clocks = sampler.stop() if sampler else None
    # with N > 1 a step contains host-side plumbing (D2H of the slice, all-gather, H2D): use the wall clock
    ms = sharding.max_over_ranks(wall if dist is not None else ms_dev, dist, device=dev)
    if rank != 0:
        return None
    peak_mac, _ = lib.bench_int_pipe(0, 20000)
    model = msm_model(n, c)
    ba = prof.get("BucketAccumulate", {"ms": 0.0, "launches": 1})
    ba_ms = ba["ms"] / max(1, ba["launches"])
    W = model["W"]
    my_windows = (W + world - 1) // world
    ba_macs = n * my_windows * 10 * MAC_PER_MODMUL
    cpu = cpu_msm_rate(min(n, 1 << 14), sample=1, procs=1)
    return {
        "metric": "g1_msm_mpoints_per_s", "value": n * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpoints/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq, 255-bit Fr)", "data": "synthetic",
        "config": {"workload": "one G1 MSM of n=%d terms (multiexp_unchecked), window c=%d, W=%d windows%s; result checked against the oracle"
                               % (n, c, W, "" if world == 1 else " split over %d GPUs + one all-gather of %d B" % (world, W * 144)),
                   "n": n, "window": c, "l2": "inputs_larger_than_l2" if n * 128 > 126 << 20 else "scratch_larger_than_l2"},
        "e2e": {"value": n * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": 0 if world == 1 else W * 144,
                "d2h_bytes_per_step": 0 if world == 1 else my_windows * 144, "note": "bases and scalars are device-resident in this workload"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "int_pipe", "kernel": "BucketAccumulate", "achieved": ba_macs / (ba_ms * 1e-3) / 1e9 if ba_ms else 0.0, "peak": peak_mac / 1e9,
                     "unit": "GMAC/s", "frac": ba_macs / (ba_ms * 1e-3) / peak_mac if ba_ms and peak_mac else None, "traffic": None,
                     "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}},
        "cpu_baseline": cpu,
    }



def run_ours(args, rank, world, dist):
    import ctypes

    from curdleproofs_pie_b200 import runtime as rt

    lib = rt.get_lib()
    assert lib.backend == "cuda-sm_100a", "bench must run on the CUDA library, got " + lib.backend
    B, n = args.batch, args.n
    c = args.window or int(lib.c.cpg_msm_pick_window_batched(B, n))
    rng = random.Random(1000 + rank)
    model = msm_model(n, c)

    # ---- synthetic inputs, resident in HBM ----
    jac = device_random_points(lib, rt, rng, B * n)
    bases = lib.jac_to_aff(jac, B * n)
    del jac
    scal_bytes = rand_scalar_bytes(rng, B * n)
    scalars = lib.upload(scal_bytes)
    out = lib.alloc(B * rt.JAC)

    # ---- host-side (pinned) copies for the e2e path ----
    comp_dev = lib.alloc(B * n * 48)
    lib.check(lib.c.cpg_g1_compress_aff(bases.ptr, B * n, comp_dev.ptr))
    h_in_pts = lib.c.cpg_host_alloc(B * n * 48)
    h_in_sc = lib.c.cpg_host_alloc(B * n * 32)
    h_out = lib.c.cpg_host_alloc(B * 48)
    lib.check(lib.c.cpg_d2h(h_in_pts, comp_dev.ptr, B * n * 48))
    ctypes.memmove(h_in_sc, scal_bytes, B * n * 32)
    e_pts = lib.alloc(B * n * 48); e_aff = lib.alloc(B * n * rt.AFF); e_err = lib.alloc(B * n)
    e_sc = lib.alloc(B * n * 32); e_out = lib.alloc(B * rt.JAC); e_comp = lib.alloc(B * 48)

    def step():
        lib.msm_batched(bases, n, scalars, B, n, c, out=out)

    def step_e2e():
        lib.check(lib.c.cpg_h2d(e_pts.ptr, h_in_pts, B * n * 48))
        lib.check(lib.c.cpg_h2d(e_sc.ptr, h_in_sc, B * n * 32))
        lib.check(lib.c.cpg_g1_decompress(e_pts.ptr, B * n, 0, e_aff.ptr, e_err.ptr))
        lib.msm_batched(e_aff, n, e_sc, B, n, c, out=e_out)
        lib.check(lib.c.cpg_g1_compress(e_out.ptr, B, e_comp.ptr))
        lib.check(lib.c.cpg_d2h(h_out, e_comp.ptr, B * 48))      # synchronises

    def barrier():
        lib.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(ms):
        from curdleproofs_pie_b200 import sharding

        return sharding.max_over_ranks(ms, dist, device="cuda")

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if rank == 0 and n <= 8192:
        # parity on this very instance: the last MSM of the batch against the oracle's C Pippenger
        from oracle import cref_binding

        cref = cref_binding.load()
        enc = lib.download(comp_dev, n * 48, offset=(B - 1) * n * 48)
        blobs = [cref.decompress(enc[48 * i:48 * i + 48], False) for i in range(n)]
        ks = [int.from_bytes(scal_bytes[32 * ((B - 1) * n + i):32 * ((B - 1) * n + i) + 32], "little") for i in range(n)]
        want = cref.compress(cref.msm(blobs, ks))
        got = lib.compress_jac(out, B)[48 * (B - 1):48 * B]
        assert got == want, "batched MSM differs from the oracle"

    # ---- timed region: K steps, device events on the launching stream, per-kernel events live ----
    sampler = ClockSampler(lib.device) if rank == 0 else None
    lib.profile(True)
    launches0 = lib.launch_count()
    barrier()
    lib.timer_start()
    for _ in range(args.steps):
        step()
    ms = lib.timer_stop()
    barrier()
    launches = lib.launch_count() - launches0
    prof = lib.profile_report()
    lib.profile(False)
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)

    # ---- e2e: host buffers in, host bytes out ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    lib.timer_start()
    for _ in range(args.steps):
        step_e2e()
    ms_e2e = lib.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel, measured live above ----
    peak_mac, _ = lib.bench_int_pipe(0, 20000)
    fq_rate, _ = lib.bench_int_pipe(2, 2000)
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    ba = prof.get("BucketAccumulate", {"ms": 0.0, "launches": 1})
    ba_ms = ba["ms"] / max(1, ba["launches"])
    ba_macs = B * model["bucket_accumulate"] * MAC_PER_MODMUL
    achieved = ba_macs / (ba_ms * 1e-3) if ba_ms else 0.0
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    roofline = {
        "bound": "int_pipe", "kernel": "BucketAccumulate", "achieved": achieved / 1e9, "peak": peak_mac / 1e9, "unit": "GMAC/s",
        "frac": achieved / peak_mac if peak_mac else None, "traffic": None,
        "peak_source": "IMAD.WIDE.U32 microbenchmark measured in this run (MEASURED_PEAKS.json has no integer-pipe entry)",
        "kernel_ms_per_launch": ba_ms, "kernel_share_of_step": ba["ms"] / total_kernel_ms if total_kernel_ms else None,
        "algorithmic_modmul_per_launch": B * model["bucket_accumulate"], "mac_per_modmul": MAC_PER_MODMUL,
        "fq_mul_chain_per_s": fq_rate,
        "whole_step_modmul_per_s": B * (model["bucket_accumulate"] + model["window_reduce"] + model["horner"]) / (ms / args.steps * 1e-3),
        "dominant_by_time": dom[0],
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
    }

    # ---- CPU baseline: oracle C port on one core, bounded sample of the same workload ----
    cpu = cpu_msm_rate(n, sample=args.cpu_sample, procs=1)
    units = B * n * world
    line = {
        "metric": "g1_msm_mpoints_per_s", "value": units * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq, 255-bit Fr)", "data": "synthetic",
        "config": {"workload": "batched G1 MSM, B=%d independent MSMs x n=%d terms per GPU (Whisk-size), window c=%d" % (B, n, c),
                   "B_per_gpu": B, "n": n, "window": c, "l2": "inputs_larger_than_l2" if B * n * 128 > 126 << 20 else "scratch_larger_than_l2",
                   "sharding": "per-MSM, no collective"},
        "msm_per_s": B * world * args.steps / (ms * 1e-3),
        "e2e": {"value": units * args.steps / (ms_e2e * 1e-3) / 1e6, "unit": "Mpoints/s",
                "h2d_bytes_per_step": B * n * 80, "d2h_bytes_per_step": B * 48, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    return line


# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    n, count, seed = job
    from oracle import cref_binding

    cref = cref_binding.load()
    rng = random.Random(seed)
    g = cref.generator()
    pts = cref.mul_batch([g] * n, [rng.randrange(1, R_ORDER) for _ in range(n)])
    rows = [[rng.randrange(R_ORDER) for _ in range(n)] for _ in range(count)]
    flat_pts = b"".join(pts)
    import ctypes

    out = ctypes.create_string_buffer(144)
    ks = [b"".join(k.to_bytes(32, "little") for k in row) for row in rows]
    t0 = time.perf_counter()
    for kb in ks:
        cref.lib.ref_g1_msm_pippenger(flat_pts, kb, n, out)
    return time.perf_counter() - t0


def cpu_msm_rate(n, sample, procs):
    """Oracle C port of arkworks' multiexp (signed... see oracle/cref) on `procs` host cores."""
    from oracle import cref_binding

    cref_binding.build()
    per = max(1, sample // procs)
    if procs == 1:
        secs = _cpu_worker((n, per, 1))
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(procs) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(n, per, i) for i in range(procs)])
            secs = time.perf_counter() - t0
    done = per * procs
    return {"value": done * n / secs / 1e6, "unit": "Mpoints/s", "cores": procs, "kind": "port",
            "sample": "%d MSMs of n=%d (Pippenger, arkworks window rule, C oracle)" % (done, n), "seconds": secs}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    n = args.n
    best = None
    t_all = 0.0
    for i in range(args.warmup + args.steps):
        r = cpu_msm_rate(n, sample=max(cores * 8, args.cpu_sample), procs=cores)
        if i >= args.warmup:
            t_all += r["seconds"]
            best = r if best is None else best
            best["value"] = max(best["value"], r["value"]) if best is not r else r["value"]
    value = best["value"]
    return {
        "impl": "reference", "metric": "g1_msm_mpoints_per_s", "value": value, "unit": "Mpoints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": {"workload": "batched G1 MSM, n=%d terms (Whisk-size); bounded sample per step" % n, "n": n},
        "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": cores, "kind": "port", "sample": best["sample"]},
        "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "py_arkworks_bls12381 (the reference's Rust arithmetic) is not installable offline; this arm times the oracle's C port of it",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="verify", choices=["verify", "prove", "msm", "msm_large"])
    ap.add_argument("--prove-window", type=int, default=0, help="bucket window of the prover's variable-base MSMs (0 = model)")
    ap.add_argument("--table-window", type=int, default=6, help="prover: window of the per-base tables of T_i / U_i multiples (0 = bucket method)")
    ap.add_argument("--prove-lanes", type=int, default=3, help="sub-batches of the prover issued alternately on separate streams")
    ap.add_argument("--fixed-window", type=int, default=16, help="window of the CRS fixed-base tables (16 = 6.6 GB table, built once per CRS; 0 = library default 12, 540 MB)")
    ap.add_argument("--prove-batch", type=int, default=4096, help="proofs in the prove side-measurement of the default (verify) run; 0 = skip")
    ap.add_argument("--cpu-sample-prove", type=int, default=4, help="proofs in the bounded CPU sample")
    ap.add_argument("--batch", type=int, default=8192, help="proofs (or MSMs) per GPU per step")
    ap.add_argument("--terms", dest="n", type=int, default=128, help="terms per MSM (msm workloads)")
    ap.add_argument("--window", type=int, default=0, help="bucket window width (0 = from the work model)")
    ap.add_argument("--transcript", default="device", choices=["device", "host"], help="where the Fiat-Shamir transcript + coefficient algebra run")
    ap.add_argument("--group", type=int, default=0, help="cross-proof aggregation: proofs per aggregated MSM (1 = per-proof MSMs, 0 = adaptive); failing groups are re-checked per proof")
    ap.add_argument("--group-window", type=int, default=0, help="window of the aggregated MSM (0 = model)")
    ap.add_argument("--corrupt-every", type=int, default=64, help="one lane in this many carries a corrupted proof or input (0 = none)")
    ap.add_argument("--streams", type=int, default=4, help="sub-batches in flight on separate CUDA streams")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads for the transcript (0 = all cores)")
    ap.add_argument("--cpu-sample", type=int, default=400, help="MSMs in the bounded CPU sample (msm workload)")
    ap.add_argument("--cpu-sample-verify", type=int, default=24, help="verifications in the bounded CPU sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = (run_reference_verify if args.workload == "verify" else run_reference)(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    os.environ["CPG_DEVICE"] = str(local)
    line = {"verify": run_ours_verify, "prove": run_ours_prove, "msm": run_ours, "msm_large": run_ours_msm_large}[args.workload](args, rank, world, dist)
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Single large MSMs (G1Point.multiexp_unchecked at n = 2^7 ... 2^20, BASELINE configs 2 and 5) on one or several
GPUs.  The W Pippenger windows are independent, so they are split across ranks; every rank holds all bases and
scalars, computes the window sums of its slice, and ONE small all-gather (W x 144 bytes in total) gives every rank
all sums for the final Horner pass.

The product path is `cpg_g1_msm_sharded`: slice, ncclAllGather on device buffers and Horner all inside libcpg.so over
the communicator of `comm.init` (no torch, nothing crosses to the host).  `gather` exists for the CPU test tier: it
replaces the transport (a gloo all-gather of byte strings in tests/test_dist_gloo.py) while the slice / combine calls
stay the library's own."""
from . import runtime as _rt
from .sharding import shard_range


def msm_large(lib, bases_aff, scalars, n, window=0, gather=None, rank=0, world=1):
    """sum_i scalars[i] * bases[i] for device-resident affine bases / canonical scalars.
    Returns a DevBuf holding the Jacobian result (identical on every rank)."""
    c = window or int(lib.c.cpg_msm_pick_window(n))
    out = lib.alloc(_rt.JAC)
    if gather is None:
        lib.check(lib.c.cpg_g1_msm_sharded(bases_aff.ptr, scalars.ptr, n, c, out.ptr), "cpg_g1_msm_sharded")
        return out
    # injected transport (tests): gather(local_bytes, width) -> [bytes per rank]
    W = int(lib.c.cpg_msm_window_count(n, c))
    lo, hi = shard_range(W, rank, world)
    local = b""
    if hi > lo:
        mine = lib.alloc((hi - lo) * _rt.JAC)
        lib.check(lib.c.cpg_g1_msm_window_sums(bases_aff.ptr, scalars.ptr, n, c, lo, hi, mine.ptr), "cpg_g1_msm_window_sums")
        local = lib.download(mine, (hi - lo) * _rt.JAC)
    width = max(h - l for l, h in (shard_range(W, r, world) for r in range(world))) * _rt.JAC
    parts = gather(local.ljust(width, b"\0"), width)
    blob = b"".join(parts[r][:(h - l) * _rt.JAC] for r, (l, h) in enumerate(shard_range(W, r, world) for r in range(world)))
    wsums = lib.upload(blob)
    lib.check(lib.c.cpg_g1_msm_combine_windows(wsums.ptr, c, out.ptr), "cpg_g1_msm_combine_windows")
    return out

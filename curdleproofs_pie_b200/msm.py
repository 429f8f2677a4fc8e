"""Single large MSMs (G1Point.multiexp_unchecked at n = 2^7 ... 2^20, BASELINE configs 2 and 5) on one
or several GPUs.  The W Pippenger windows are independent, so they are split across ranks; every rank
holds all bases and scalars, computes the window sums of its slice, and ONE small all-gather
(W x 144 bytes in total) gives every rank all sums for the final Horner pass.
torch.distributed is only the transport of those few KB (NCCL over NVLink on the GPU box, gloo in tests)."""
from . import runtime as _rt
from .sharding import shard_range


def msm_large(lib, bases_aff, scalars, n, window=0, dist=None, device="cpu"):
    """sum_i scalars[i] * bases[i] for device-resident affine bases / canonical scalars.
    Returns a DevBuf holding the Jacobian result (identical on every rank)."""
    c = window or int(lib.c.cpg_msm_pick_window(n))
    W = int(lib.c.cpg_msm_window_count(n, c))
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    out = lib.alloc(_rt.JAC)
    if world == 1:
        lib.check(lib.c.cpg_g1_msm_batched(bases_aff.ptr, 0, scalars.ptr, 1, n, c, out.ptr), "cpg_g1_msm_batched")
        return out
    lo, hi = shard_range(W, rank, world)
    wsums = lib.alloc(W * _rt.JAC)
    if hi > lo:
        mine = lib.alloc((hi - lo) * _rt.JAC)
        lib.check(lib.c.cpg_g1_msm_window_sums(bases_aff.ptr, scalars.ptr, n, c, lo, hi, mine.ptr), "cpg_g1_msm_window_sums")
        local = lib.download(mine, (hi - lo) * _rt.JAC)
    else:
        local = b""
    # the one collective of the path: all-gather of the per-rank window sums
    import torch

    width = max(h - l for l, h in (shard_range(W, r, world) for r in range(world))) * _rt.JAC
    buf = torch.zeros(width, dtype=torch.uint8, device=device)
    if local:
        buf[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(device)
    gathered = [torch.zeros(width, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(gathered, buf)
    parts = []
    for r in range(world):
        l, h = shard_range(W, r, world)
        parts.append(bytes(gathered[r][:(h - l) * _rt.JAC].cpu().numpy().tobytes()))
    lib.upload(b"".join(parts), wsums)
    lib.check(lib.c.cpg_g1_msm_combine_windows(wsums.ptr, c, out.ptr), "cpg_g1_msm_combine_windows")
    return out

"""Per-proof sharding across GPUs: proofs are independent, so rank r of W owns a
contiguous block of the batch and no data-path collective is needed; only the verdict bytes are
gathered at the end.  torch.distributed is plumbing here (NCCL on the GPU box, gloo in CPU tests)."""


def shard_range(total, rank, world):
    """Contiguous [lo, hi) block of `total` items for `rank`; blocks differ in size by at most 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, dist=None, device="cpu"):
    """Max of a per-rank scalar (device time in ms) over all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_verdicts(local, total, dist=None, device="cpu"):
    """Concatenate every rank's verdict bytes (its shard_range block) into the full bitmap, on all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(local)
    import torch

    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=torch.uint8, device=device)
    if len(local):
        buf[:len(local)] = torch.tensor(list(local), dtype=torch.uint8, device=device)
    outs = [torch.zeros(width, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return b"".join(bytes(outs[r][:hi - lo].cpu().tolist()) for r, (lo, hi) in enumerate(sizes))

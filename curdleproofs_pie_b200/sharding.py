"""Per-proof sharding across GPUs: proofs are independent, so rank r of W owns a contiguous block of
the batch and no data-path collective is needed; only the verdict bytes (and a timing scalar) are gathered at the end.
The transport is the library's own communicator (comm.py: NCCL inside libcpg.so); callers without one - the CPU test
tier - inject `allgather(local_bytes, width) -> [bytes per rank]`.  No torch here."""
import struct


def shard_range(total, rank, world):
    """Contiguous [lo, hi) block of `total` items for `rank`; blocks differ in size by at most 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _transport(lib, allgather):
    if allgather is not None:
        return allgather
    from . import comm

    return lambda local, width: comm.allgather_bytes(lib, local, width)


def max_over_ranks(value, lib=None, allgather=None, world=None):
    """Max of a per-rank scalar (device time in ms) over all ranks."""
    if world is None:
        world = int(lib.c.cpg_comm_world()) if lib is not None else 1
    if world == 1 and allgather is None:
        return float(value)
    parts = _transport(lib, allgather)(struct.pack("<d", float(value)), 8)
    return max(struct.unpack("<d", p[:8])[0] for p in parts)


def gather_verdicts(local, total, rank=0, world=1, lib=None, allgather=None):
    """Concatenate every rank's verdict bytes (its shard_range block) into the full bitmap, on all ranks."""
    if world == 1:
        return bytes(local)
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    parts = _transport(lib, allgather)(bytes(local).ljust(width, b"\0"), width)
    return b"".join(parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes))

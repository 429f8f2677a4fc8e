"""The N>1 path on CPU: world_size-2 gloo processes exercise the per-proof sharding, the max-over-ranks timing
reduction, the verdict gather and the window-split MSM.  On the GPU box the transport is the library's own NCCL
communicator (comm.py); here a gloo all-gather of byte strings is injected in its place (the product package itself
holds no torch import)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def gloo_allgather(dist):
    """allgather(local_bytes, width) -> [bytes per rank] over torch.distributed (gloo)"""
    import torch

    def gather(local, width):
        buf = torch.frombuffer(bytearray(bytes(local).ljust(width, b"\0")), dtype=torch.uint8)
        outs = [torch.zeros(width, dtype=torch.uint8) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, buf)
        return [bytes(o.numpy().tobytes()) for o in outs]

    return gather


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from curdleproofs_pie_b200 import sharding
    from test_dist_gloo import gloo_allgather

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(total, rank, world)
    local = bytes((i * 7 + 3) % 2 for i in range(lo, hi))           # this rank's "verdicts"
    ag = gloo_allgather(dist)
    full = sharding.gather_verdicts(local, total, rank, world, allgather=ag)
    slowest = sharding.max_over_ranks(10.0 + rank, allgather=ag, world=world)
    dist.barrier()
    q.put((rank, lo, hi, full, slowest))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7, 1])
def test_two_rank_sharding_and_gather(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes((i * 7 + 3) % 2 for i in range(total))
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total        # contiguous cover
    for rank, lo, hi, full, slowest in res:
        assert full == want
        assert slowest == 11.0


def test_shard_range_properties():
    from curdleproofs_pie_b200.sharding import shard_range

    for total in (0, 1, 5, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _msm_worker(rank, world, port, n, q):
    """Window-split large MSM: each rank computes its slice of Pippenger windows on the host-emulated
    kernels, one gloo all-gather exchanges the window sums, every rank combines."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import random

    import torch.distributed as dist

    import conftest
    import parity_cases as pc
    from curdleproofs_pie_b200 import msm, runtime
    from test_dist_gloo import gloo_allgather
    from oracle import cref_binding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = runtime.CpgLib(conftest.SEAM_SO, 0)
    cref = cref_binding.load()
    rng = random.Random(31)                       # same inputs on every rank
    blobs, enc = pc.rand_points(cref, rng, 16)
    idx = [rng.randrange(16) for _ in range(n)]
    ks = [rng.randrange(pc.R) for _ in range(n)]
    aff, _ = pc.upload_points(lib, [enc[i] for i in idx])
    dk = lib.upload(runtime.scalars_to_bytes(ks))
    out = msm.msm_large(lib, aff, dk, n, window=6, gather=gloo_allgather(dist), rank=rank, world=world)
    # world 1 in the library (no communicator): the sharded entry point is the plain MSM
    solo = msm.msm_large(lib, aff, dk, n, window=6)
    assert lib.compress_jac(solo, 1) == lib.compress_jac(out, 1)
    assert lib.c.cpg_comm_world() == 1 and lib.c.cpg_comm_rank() == 0
    got = lib.compress_jac(out, 1)
    agg = [0] * 16
    for i, k in zip(idx, ks):
        agg[i] = (agg[i] + k) % pc.R
    want = cref.compress(cref.msm(blobs, agg))
    dist.barrier()
    q.put((rank, got == want))
    dist.destroy_process_group()


def test_two_rank_window_split_msm(seam_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_msm_worker, args=(r, 2, port, 200, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]

"""Host plumbing of the library's communicator (include/cpg.h: cpg_comm_*): one process per GPU, NCCL inside
libcpg.so.  The only thing the host has to do is carry rank 0's 128-byte id to the other ranks before
cpg_comm_init; this module does that over a TCP socket on MASTER_ADDR (standard library only - no torch), or through
any callable the caller provides.

    comm.init_from_env(lib)        # RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as set by torchrun
    lib.c.cpg_g1_msm_sharded(...)  # collectives now run over NVLink on device buffers
"""
import ctypes
import os
import socket
import time

ID_BYTES = 128
PORT_OFFSET = 29        # the id is served on MASTER_PORT + 29 (MASTER_PORT itself belongs to the launcher's store)


def _serve_id(addr, port, blob, nclients, timeout):
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(nclients)
        srv.settimeout(timeout)
        for _ in range(nclients):
            conn, _peer = srv.accept()
            with conn:
                conn.sendall(blob)


def _fetch_id(addr, port, timeout):
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5) as s:
                blob = b""
                while len(blob) < ID_BYTES:
                    part = s.recv(ID_BYTES - len(blob))
                    if not part:
                        break
                    blob += part
                if len(blob) == ID_BYTES:
                    return blob
        except OSError:
            pass
        if time.monotonic() > deadline:
            raise TimeoutError("no communicator id from rank 0 at %s:%d" % (addr, port))
        time.sleep(0.05)


def init(lib, rank, world, exchange=None, addr="127.0.0.1", port=29529, timeout=120.0):
    """Create the communicator on `lib`'s device.  exchange(blob_or_None) -> blob lets the caller move the id itself
    (rank 0 passes the id in and gets it back, the others pass None); default: TCP from rank 0."""
    rank, world = int(rank), int(world)
    if world == 1:
        lib.check(lib.c.cpg_comm_init(0, 1, None), "cpg_comm_init")
        return
    blob = None
    if rank == 0:
        buf = ctypes.create_string_buffer(ID_BYTES)
        lib.check(lib.c.cpg_comm_unique_id(buf), "cpg_comm_unique_id")
        blob = buf.raw
    if exchange is not None:
        blob = exchange(blob)
    elif rank == 0:
        _serve_id(addr, port, blob, world - 1, timeout)
    else:
        blob = _fetch_id(addr, port, timeout)
    if not blob or len(blob) != ID_BYTES:
        raise ValueError("communicator id must be %d bytes" % ID_BYTES)
    lib.check(lib.c.cpg_comm_init(rank, world, blob), "cpg_comm_init")


def init_from_env(lib, timeout=120.0):
    """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as torchrun (or any launcher) sets them."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + PORT_OFFSET
    init(lib, rank, world, addr=addr, port=port, timeout=timeout)
    return rank, world


def allgather_bytes(lib, local, width=None):
    """All-gather of one equal-length byte string per rank through the library (device buffers, NCCL)."""
    world = int(lib.c.cpg_comm_world())
    local = bytes(local)
    width = len(local) if width is None else width
    if world == 1:
        return [local]
    send = lib.upload(local.ljust(width, b"\0"))
    recv = lib.alloc(world * width)
    lib.check(lib.c.cpg_comm_allgather(send.ptr, recv.ptr, width), "cpg_comm_allgather")
    raw = lib.download(recv, world * width)
    return [raw[r * width:(r + 1) * width] for r in range(world)]

"""ctypes binding of libcpg.so (include/cpg.h).  No torch, no Triton, no CPU fallback.

``get_lib()`` loads the nvcc-built ``curdleproofs_pie_b200/lib/libcpg.so`` and initialises the
CUDA device; it raises ``RuntimeError`` when the library has not been built or no B200 is
visible.  Nothing in this package computes group arithmetic on the host.

(The CPU-only test tier exercises the same launch logic through a host emulation of the
kernels, built by tests/conftest.py into tests/_build/; it installs it with
``_install_library_for_tests`` - the product never looks for it.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcpg.so")

AFF = 96
JAC = 144
SCALAR = 32
COMPRESSED = 48

R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

_c = ctypes
_SIGS = {
    "cpg_init": (_c.c_int, [_c.c_int]),
    "cpg_device_count": (_c.c_int, []),
    "cpg_last_error": (_c.c_char_p, []),
    "cpg_backend": (_c.c_char_p, []),
    "cpg_set_stream": (_c.c_int, [_c.c_void_p]),
    "cpg_sync": (_c.c_int, []),
    "cpg_malloc": (_c.c_void_p, [_c.c_size_t]),
    "cpg_free": (_c.c_int, [_c.c_void_p]),
    "cpg_memset": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_size_t]),
    "cpg_h2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t]),
    "cpg_d2h": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t]),
    "cpg_d2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t]),
    "cpg_host_alloc": (_c.c_void_p, [_c.c_size_t]),
    "cpg_host_free": (_c.c_int, [_c.c_void_p]),
    "cpg_timer_start": (_c.c_int, []),
    "cpg_timer_stop": (_c.c_int, [_c.POINTER(_c.c_float)]),
    "cpg_launch_count": (_c.c_uint64, []),
    "cpg_profile_enable": (_c.c_int, [_c.c_int]),
    "cpg_profile_reset": (_c.c_int, []),
    "cpg_profile_report": (_c.c_int, [_c.c_char_p, _c.c_size_t]),
    "cpg_g1_decompress": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "cpg_g1_compress": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_compress_aff": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_aff_to_jac": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_jac_to_aff": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_generator": (_c.c_int, [_c.c_void_p]),
    "cpg_g1_identity": (_c.c_int, [_c.c_void_p]),
    "cpg_g1_add": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_sub": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_neg": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_eq": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_is_identity": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_mul": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_fold": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_void_p]),
    "cpg_g1_msm_batched": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "cpg_g1_msm_batched_off": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "cpg_msm_pick_window": (_c.c_int, [_c.c_size_t]),
    "cpg_msm_force_path": (_c.c_int, [_c.c_int]),
    "cpg_msm_set_accumulate": (_c.c_int, [_c.c_int]),
    "cpg_msm_pick_window_batched": (_c.c_int, [_c.c_size_t, _c.c_size_t]),
    "cpg_msm_window_count": (_c.c_int, [_c.c_size_t, _c.c_int]),
    "cpg_g1_msm_window_sums": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "cpg_g1_msm_combine_windows": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p]),
    "cpg_comm_unique_id": (_c.c_int, [_c.c_char_p]),
    "cpg_comm_init": (_c.c_int, [_c.c_int, _c.c_int, _c.c_char_p]),
    "cpg_comm_free": (_c.c_int, []),
    "cpg_comm_rank": (_c.c_int, []),
    "cpg_comm_world": (_c.c_int, []),
    "cpg_comm_nccl_version": (_c.c_int, []),
    "cpg_comm_allgather": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t]),
    "cpg_g1_msm_sharded": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "cpg_fixed_table_create": (_c.c_void_p, [_c.c_void_p, _c.c_size_t, _c.c_int]),
    "cpg_fixed_table_free": (_c.c_int, [_c.c_void_p]),
    "cpg_fixed_table_bytes": (_c.c_size_t, [_c.c_void_p]),
    "cpg_g1_msm_fixed_batched": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "cpg_fr_add": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_fr_sub": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_fr_mul": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_fr_inverse": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "cpg_merlin_script": (_c.c_int, [_c.c_char_p, _c.c_size_t, _c.c_int, _c.c_char_p, _c.c_size_t, _c.POINTER(_c.c_size_t)]),
    "cpg_merlin_new": (_c.c_void_p, [_c.c_char_p, _c.c_size_t]),
    "cpg_merlin_clone": (_c.c_void_p, [_c.c_void_p]),
    "cpg_merlin_free": (_c.c_int, [_c.c_void_p]),
    "cpg_merlin_append": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t, _c.c_char_p, _c.c_size_t]),
    "cpg_merlin_challenge": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t, _c.c_char_p, _c.c_size_t]),
    "cpg_verifier_create": (_c.c_void_p, [_c.c_char_p, _c.c_size_t, _c.c_size_t, _c.c_int, _c.c_int]),
    "cpg_verifier_create_sharded": (_c.c_void_p, [_c.c_char_p, _c.c_size_t, _c.c_size_t, _c.c_int, _c.c_int]),
    "cpg_prover_create_sharded": (_c.c_void_p, [_c.c_char_p, _c.c_size_t, _c.c_size_t, _c.c_int]),
    "cpg_verifier_free": (_c.c_int, [_c.c_void_p]),
    "cpg_verifier_proof_bytes": (_c.c_size_t, [_c.c_void_p]),
    "cpg_verifier_input_bytes": (_c.c_size_t, [_c.c_void_p]),
    "cpg_verifier_set_window": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_verifier_set_transcript": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_verifier_set_streams": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_verifier_set_group": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int]),
    "cpg_verifier_set_cache": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_verifier_cache_reset": (_c.c_int, [_c.c_void_p]),
    "cpg_verifier_cache_stats": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_uint64)]),
    "cpg_verifier_rechecked": (_c.c_size_t, [_c.c_void_p]),
    "cpg_verifier_group": (_c.c_int, [_c.c_void_p]),
    "cpg_verify_batch": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_char_p, _c.c_size_t, _c.c_char_p]),
    "cpg_verify_replay_device": (_c.c_int, [_c.c_void_p, _c.c_char_p]),
    "cpg_prover_create": (_c.c_void_p, [_c.c_char_p, _c.c_size_t, _c.c_size_t, _c.c_int]),
    "cpg_prover_free": (_c.c_int, [_c.c_void_p]),
    "cpg_prover_proof_bytes": (_c.c_size_t, [_c.c_void_p]),
    "cpg_prover_rand_scalars": (_c.c_size_t, [_c.c_void_p]),
    "cpg_prover_set_window": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_pyrandom_draw_shuffles": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_size_t, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "cpg_prover_set_table_window": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_prover_set_transcript": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "cpg_prover_set_lanes": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_size_t]),
    "cpg_prove_replay_device": (_c.c_int, [_c.c_void_p]),
    "cpg_prove_batch": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.c_char_p, _c.c_char_p, _c.c_size_t, _c.c_char_p, _c.c_char_p, _c.c_char_p]),
    "cpg_bench_int_pipe": (_c.c_int, [_c.c_int, _c.c_uint64, _c.POINTER(_c.c_double), _c.POINTER(_c.c_float)]),
}
EXPORTS = tuple(sorted(_SIGS))


class CpgError(RuntimeError):
    pass


class DevBuf:
    """Owning handle of one device allocation."""

    __slots__ = ("lib", "ptr", "nbytes")

    def __init__(self, lib, nbytes):
        self.lib = lib
        self.nbytes = int(nbytes)
        self.ptr = lib.c.cpg_malloc(max(1, self.nbytes))
        if not self.ptr:
            raise CpgError("cpg_malloc(%d) failed: %s" % (nbytes, lib.last_error()))

    def free(self):
        if self.ptr:
            self.lib.c.cpg_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CpgLib:
    def __init__(self, path=LIB_PATH, device=0):
        if not os.path.exists(path):
            raise CpgError(
                "CUDA library %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % path
            )
        self.path = path
        self.c = ctypes.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.c, name)
            fn.restype = res
            fn.argtypes = args
        rc = self.c.cpg_init(int(device))
        if rc:
            raise CpgError("cpg_init(%d) failed: %s" % (device, self.last_error()))
        self.device = device
        self.backend = self.c.cpg_backend().decode()

    # -- plumbing --
    def last_error(self):
        return (self.c.cpg_last_error() or b"").decode()

    def check(self, rc, what=""):
        if rc:
            raise CpgError("%s failed: %s" % (what or "libcpg call", self.last_error()))

    def alloc(self, nbytes):
        return DevBuf(self, nbytes)

    def upload(self, data, buf=None):
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        if buf is None:
            buf = self.alloc(len(data))
        if len(data):
            src = (ctypes.c_char * len(data)).from_buffer_copy(data)
            self.check(self.c.cpg_h2d(buf.ptr, src, len(data)), "cpg_h2d")
            self.check(self.c.cpg_sync(), "cpg_sync")  # src is a temporary
        return buf

    def download(self, buf, nbytes=None, offset=0):
        nbytes = buf.nbytes - offset if nbytes is None else nbytes
        out = ctypes.create_string_buffer(max(1, nbytes))
        if nbytes:
            self.check(self.c.cpg_d2h(out, buf.ptr + offset, nbytes), "cpg_d2h")
        return out.raw[:nbytes]

    def sync(self):
        self.check(self.c.cpg_sync(), "cpg_sync")

    def launch_count(self):
        return int(self.c.cpg_launch_count())

    # -- serialisation --
    def decompress(self, data48, check_subgroup=False):
        """bytes (k*48) -> (DevBuf of k affine points, list of per-point error codes)."""
        k = len(data48) // COMPRESSED
        src = self.upload(data48)
        out = self.alloc(k * AFF)
        err = self.alloc(k)
        self.check(self.c.cpg_g1_decompress(src.ptr, k, 1 if check_subgroup else 0, out.ptr, err.ptr), "cpg_g1_decompress")
        return out, list(self.download(err, k))

    def compress_jac(self, jac, k):
        out = self.alloc(k * COMPRESSED)
        self.check(self.c.cpg_g1_compress(jac.ptr, k, out.ptr), "cpg_g1_compress")
        return self.download(out, k * COMPRESSED)

    def compress_aff(self, aff, k):
        out = self.alloc(k * COMPRESSED)
        self.check(self.c.cpg_g1_compress_aff(aff.ptr, k, out.ptr), "cpg_g1_compress_aff")
        return self.download(out, k * COMPRESSED)

    def aff_to_jac(self, aff, k):
        out = self.alloc(k * JAC)
        self.check(self.c.cpg_g1_aff_to_jac(aff.ptr, k, out.ptr), "cpg_g1_aff_to_jac")
        return out

    def jac_to_aff(self, jac, k):
        out = self.alloc(k * AFF)
        self.check(self.c.cpg_g1_jac_to_aff(jac.ptr, k, out.ptr), "cpg_g1_jac_to_aff")
        return out

    def generator(self):
        out = self.alloc(JAC)
        self.check(self.c.cpg_g1_generator(out.ptr), "cpg_g1_generator")
        return out

    def identity(self):
        out = self.alloc(JAC)
        self.check(self.c.cpg_g1_identity(out.ptr), "cpg_g1_identity")
        return out

    # -- group law --
    def add(self, a, b, k):
        out = self.alloc(k * JAC)
        self.check(self.c.cpg_g1_add(a.ptr, b.ptr, k, out.ptr), "cpg_g1_add")
        return out

    def sub(self, a, b, k):
        out = self.alloc(k * JAC)
        self.check(self.c.cpg_g1_sub(a.ptr, b.ptr, k, out.ptr), "cpg_g1_sub")
        return out

    def neg(self, a, k):
        out = self.alloc(k * JAC)
        self.check(self.c.cpg_g1_neg(a.ptr, k, out.ptr), "cpg_g1_neg")
        return out

    def eq(self, a, b, k):
        out = self.alloc(k)
        self.check(self.c.cpg_g1_eq(a.ptr, b.ptr, k, out.ptr), "cpg_g1_eq")
        return list(self.download(out, k))

    def is_identity(self, a, k):
        out = self.alloc(k)
        self.check(self.c.cpg_g1_is_identity(a.ptr, k, out.ptr), "cpg_g1_is_identity")
        return list(self.download(out, k))

    def mul(self, p, scalars, k, group=1):
        out = self.alloc(k * JAC)
        self.check(self.c.cpg_g1_mul(p.ptr, scalars.ptr, k, group, out.ptr), "cpg_g1_mul")
        return out

    def fold(self, L, R, x, rows, m):
        out = self.alloc(rows * m * JAC)
        self.check(self.c.cpg_g1_fold(L.ptr, R.ptr, x.ptr, rows, m, out.ptr), "cpg_g1_fold")
        return out

    # -- MSM --
    def msm_batched(self, bases_aff, base_stride, scalars, B, n, window=0, out=None):
        if out is None:
            out = self.alloc(B * JAC)
        self.check(self.c.cpg_g1_msm_batched(bases_aff.ptr, base_stride, scalars.ptr, B, n, window, out.ptr), "cpg_g1_msm_batched")
        return out

    def fixed_table(self, bases_aff, nb, window=0):
        t = self.c.cpg_fixed_table_create(bases_aff.ptr, nb, window)
        if not t:
            raise CpgError("cpg_fixed_table_create failed: " + self.last_error())
        return FixedTable(self, t, nb)

    def msm_fixed_batched(self, table, scalars, B, accumulate=False, out=None):
        if out is None:
            if accumulate:
                raise ValueError("accumulate needs an output buffer")
            out = self.alloc(B * JAC)
        self.check(self.c.cpg_g1_msm_fixed_batched(table.handle, scalars.ptr, B, 1 if accumulate else 0, out.ptr), "cpg_g1_msm_fixed_batched")
        return out

    # -- Fr --
    def fr_op(self, name, a, b, k):
        out = self.alloc(k * SCALAR)
        fn = getattr(self.c, "cpg_fr_" + name)
        if name == "inverse":
            self.check(fn(a.ptr, k, out.ptr), "cpg_fr_inverse")
        else:
            self.check(fn(a.ptr, b.ptr, k, out.ptr), "cpg_fr_" + name)
        return out

    # -- transcript --
    def merlin_script(self, records, on_device=False, cap=1 << 16):
        """records: (op, more, label, data_or_n) tuples (include/cpg.h: cpg_merlin_script) -> concatenated outputs"""
        blob = bytearray()
        for op, more, label, data in records:
            emits = op in (3, 7)
            n = int(data) if emits else len(data)
            blob += bytes([op, 1 if more else 0]) + len(label).to_bytes(2, "little") + n.to_bytes(4, "little") + bytes(label)
            if not emits:
                blob += bytes(data)
        out = ctypes.create_string_buffer(cap)
        got = ctypes.c_size_t()
        self.check(self.c.cpg_merlin_script(bytes(blob), len(blob), int(on_device), out, cap, ctypes.byref(got)), "cpg_merlin_script")
        return out.raw[:got.value]

    def bench_int_pipe(self, kind, iters):
        per_s = ctypes.c_double()
        ms = ctypes.c_float()
        self.check(self.c.cpg_bench_int_pipe(kind, iters, ctypes.byref(per_s), ctypes.byref(ms)), "cpg_bench_int_pipe")
        return per_s.value, ms.value

    def profile(self, on):
        self.check(self.c.cpg_profile_reset(), "cpg_profile_reset")
        self.check(self.c.cpg_profile_enable(1 if on else 0), "cpg_profile_enable")

    def profile_report(self):
        import json

        buf = ctypes.create_string_buffer(1 << 16)
        self.check(self.c.cpg_profile_report(buf, len(buf)), "cpg_profile_report")
        rep = json.loads(buf.value.decode() or "{}")
        # non-zero coefficients the table-lookup MSMs met while profiling (their work model), kept beside the kernel times
        self.last_counters = rep.pop("_counters", {})
        return rep

    def timer_start(self):
        self.check(self.c.cpg_timer_start(), "cpg_timer_start")

    def timer_stop(self):
        ms = ctypes.c_float()
        self.check(self.c.cpg_timer_stop(ctypes.byref(ms)), "cpg_timer_stop")
        return ms.value


class FixedTable:
    def __init__(self, lib, handle, nb):
        self.lib = lib
        self.handle = handle
        self.nb = nb

    @property
    def nbytes(self):
        return int(self.lib.c.cpg_fixed_table_bytes(self.handle))

    def free(self):
        if self.handle:
            self.lib.c.cpg_fixed_table_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def scalars_to_bytes(values):
    return b"".join((int(v) % R_ORDER).to_bytes(32, "little") for v in values)


_LIB = None


def get_lib():
    """The process-wide library instance (one process per GPU: device = LOCAL_RANK or 0)."""
    global _LIB
    if _LIB is None:
        dev = int(os.environ.get("CPG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _LIB = CpgLib(LIB_PATH, dev)
    return _LIB


def _install_library_for_tests(lib):
    """Test seam only (tests/conftest.py): make get_lib() return an already-constructed CpgLib."""
    global _LIB
    _LIB = lib

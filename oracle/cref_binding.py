"""ORACLE (test infrastructure): ctypes binding of oracle/cref/libbls12381_ref.so."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "cref", "libbls12381_ref.so")
_INSTANCE = None


def build(force=False):
    src = os.path.join(_HERE, "cref", "bls12381_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "cref")])
    return _SO


class CRef:
    def __init__(self):
        self.lib = ctypes.CDLL(build())
        L = self.lib
        self._buf = ctypes.create_string_buffer
        L.ref_g1_eq.restype = ctypes.c_int
        L.ref_g1_decompress.restype = ctypes.c_int
        for f in ("ref_g1_msm_naive", "ref_g1_msm_pippenger", "ref_g1_mul_batch"):
            getattr(L, f).argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.ref_g1_compress_batch.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.ref_g1_decompress_batch.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]

    def generator(self):
        o = self._buf(144); self.lib.ref_g1_generator(o); return o.raw

    def identity(self):
        o = self._buf(144); self.lib.ref_g1_identity(o); return o.raw

    def add(self, a, b):
        o = self._buf(144); self.lib.ref_g1_add(a, b, o); return o.raw

    def sub(self, a, b):
        o = self._buf(144); self.lib.ref_g1_sub(a, b, o); return o.raw

    def neg(self, a):
        o = self._buf(144); self.lib.ref_g1_neg(a, o); return o.raw

    def mul(self, a, k):
        o = self._buf(144); self.lib.ref_g1_mul(a, int(k).to_bytes(32, "little"), o); return o.raw

    def eq(self, a, b):
        return bool(self.lib.ref_g1_eq(a, b))

    def compress(self, a):
        o = self._buf(48); self.lib.ref_g1_compress(a, o); return o.raw

    def decompress(self, data, check):
        if len(data) != 48:
            raise ValueError("compressed G1 point must be 48 bytes")
        o = self._buf(144)
        rc = self.lib.ref_g1_decompress(data, 1 if check else 0, o)
        if rc:
            raise ValueError("invalid compressed G1 point (code %d)" % rc)
        return o.raw

    def msm(self, pts, ks, naive=False):
        n = len(pts)
        o = self._buf(144)
        f = self.lib.ref_g1_msm_naive if naive else self.lib.ref_g1_msm_pippenger
        f(b"".join(pts), b"".join(int(k).to_bytes(32, "little") for k in ks), n, o)
        return o.raw

    def mul_batch(self, pts, ks):
        n = len(pts)
        o = self._buf(144 * n)
        self.lib.ref_g1_mul_batch(b"".join(pts), b"".join(int(k).to_bytes(32, "little") for k in ks), n, o)
        return [o.raw[144 * i:144 * (i + 1)] for i in range(n)]

    def compress_batch(self, pts):
        n = len(pts)
        o = self._buf(48 * n)
        self.lib.ref_g1_compress_batch(b"".join(pts), n, o)
        return [o.raw[48 * i:48 * (i + 1)] for i in range(n)]

    def decompress_batch(self, datas, check=False):
        n = len(datas)
        o = self._buf(144 * n)
        ok = self._buf(n)
        self.lib.ref_g1_decompress_batch(b"".join(datas), n, 1 if check else 0, o, ok)
        return [o.raw[144 * i:144 * (i + 1)] for i in range(n)], list(ok.raw)

    def keccak_f1600(self, state):
        b = self._buf(bytes(state), 200)
        self.lib.ref_keccak_f1600(b)
        return bytearray(b.raw)


def load():
    global _INSTANCE
    if _INSTANCE is None:
        _INSTANCE = CRef()
    return _INSTANCE

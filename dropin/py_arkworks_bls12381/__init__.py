"""Drop-in ``py_arkworks_bls12381`` backed by the B200 kernels (libcpg.so via ctypes).

Put ``dropin/`` ahead of site-packages on ``sys.path`` and the unmodified reference
(``curdleproofs``: CurdleProofsProof.new/verify, the Whisk API) runs on the GPU path:
    PYTHONPATH=/root/repo/dropin:/root/repo  python -m pytest curdleproofs/

Surface: /root/reference/curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-54 plus the
extra dunders pinned by curdleproofs/curdleproofs/test_curdleproofs.py:45-128.

How a G1Point is held.  The reference's ``compute_MSM`` is a Python loop
``acc = acc + base * scalar`` (cp/msm_accumulator.py:6-12), one group operation per call.  Here
``*`` and ``+`` are *lazy*: a point is a formal linear combination  sum_i k_i * L_i  over
concrete device-format leaves, and nothing is computed until a value is observed
(``to_compressed_bytes``, ``==``, ``str``).  Observation evaluates the whole combination as ONE
batched-Pippenger launch sequence on the GPU (cpg_g1_msm_batched), so every ``compute_MSM`` loop,
every IPA/SameMSM fold and every commitment of the reference becomes a real MSM without touching
the reference's code.  Group elements are never computed on the host: the only host arithmetic
is the Fr bookkeeping of the coefficients k_i (Python ints mod r), as SURVEY 8b assigns.
Results are canonical (affine, fully reduced), so bytes and equality match arkworks exactly.
"""
from curdleproofs_pie_b200 import runtime as _rt

_R = _rt.R_ORDER
_ZERO_AFF = bytes(_rt.AFF)

__all__ = ["G1Point", "Scalar"]


class Scalar:
    """Fr element; canonical value kept on the host (stub :32-54)."""

    __slots__ = ("v",)

    def __init__(self, value=0):
        if isinstance(value, Scalar):
            value = value.v
        if isinstance(value, bool) or not isinstance(value, int):
            raise TypeError("argument 'integer': 'int' expected")
        if value < 0:
            raise OverflowError("can't convert negative int to unsigned")
        self.v = value % _R

    @staticmethod
    def _raw(v):
        s = Scalar.__new__(Scalar)
        s.v = v
        return s

    def __add__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((self.v + o.v) % _R)

    __radd__ = __add__

    def __sub__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((self.v - o.v) % _R)

    def __rsub__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((o.v - self.v) % _R)

    def __mul__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw(self.v * o.v % _R)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return self * o.inverse()

    def __rtruediv__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return o * self.inverse()

    def __neg__(self):
        return Scalar._raw(-self.v % _R)

    def __eq__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return self.v == o.v

    def __ne__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return self.v != o.v

    __hash__ = None

    def __int__(self):
        return self.v

    def __str__(self):
        return self.v.to_bytes(32, "little").hex()

    __repr__ = __str__

    def inverse(self):
        # zero maps to zero so that cp/util.py:51-54's own assert is what fires
        return Scalar._raw(pow(self.v, -1, _R) if self.v else 0)

    def square(self):
        return Scalar._raw(self.v * self.v % _R)

    def pow(self, e):
        return Scalar._raw(pow(self.v, int(e), _R))

    def is_zero(self):
        return self.v == 0

    def to_le_bytes(self):
        return self.v.to_bytes(32, "little")

    @staticmethod
    def from_le_bytes(data):
        data = bytes(data)
        if len(data) != 32:
            raise ValueError("serialised data seems to be invalid")
        v = int.from_bytes(data, "little")
        if v >= _R:
            raise ValueError("serialised data seems to be invalid")
        return Scalar._raw(v)


class G1Point:
    """BLS12-381 G1 element (stub :5-30).  ``G1Point()`` is the generator."""

    __slots__ = ("_aff", "_terms", "_comp")

    def __init__(self):
        lib = _rt.get_lib()
        aff = lib.jac_to_aff(lib.generator(), 1)
        self._aff = lib.download(aff, _rt.AFF)
        self._terms = None
        self._comp = None

    # -- construction helpers --
    @staticmethod
    def _concrete(aff_bytes):
        p = G1Point.__new__(G1Point)
        p._aff = aff_bytes
        p._terms = None
        p._comp = None
        return p

    @staticmethod
    def _lazy(terms):
        p = G1Point.__new__(G1Point)
        p._aff = None
        p._terms = terms
        p._comp = None
        return p

    @staticmethod
    def identity():
        return G1Point._concrete(_ZERO_AFF)

    def _as_terms(self):
        """{id(leaf): (leaf, coefficient)} view of this point."""
        if self._aff is not None:
            return {} if self._aff == _ZERO_AFF else {id(self): (self, 1)}
        return self._terms

    # -- lazy group law --
    def __add__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        a, b = self._as_terms(), o._as_terms()
        if len(a) < len(b):
            a, b = b, a
        out = dict(a)
        for key, (leaf, k) in b.items():
            cur = out.get(key)
            if cur is None:
                out[key] = (leaf, k)
            else:
                s = (cur[1] + k) % _R
                if s:
                    out[key] = (leaf, s)
                else:
                    del out[key]
        return G1Point._lazy(out)

    __radd__ = __add__

    def __neg__(self):
        return G1Point._lazy({key: (leaf, _R - k) for key, (leaf, k) in self._as_terms().items()})

    def __sub__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        return self + (-o)

    def __rsub__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        return o + (-self)

    def __mul__(self, s):
        if not isinstance(s, Scalar):
            return NotImplemented
        v = s.v
        if v == 0:
            return G1Point._lazy({})
        return G1Point._lazy({key: (leaf, k * v % _R) for key, (leaf, k) in self._as_terms().items()})

    __rmul__ = __mul__

    # -- observation: one GPU MSM --
    def _force(self):
        if self._aff is not None:
            return self._aff
        terms = self._terms
        n = len(terms)
        if n == 0:
            self._aff = _ZERO_AFF
        else:
            lib = _rt.get_lib()
            leaves = list(terms.values())
            bases = lib.upload(b"".join(leaf._aff for leaf, _ in leaves))
            scalars = lib.upload(b"".join(k.to_bytes(32, "little") for _, k in leaves))
            jac = lib.msm_batched(bases, 0, scalars, 1, n)
            self._aff = lib.download(lib.jac_to_aff(jac, 1), _rt.AFF)
        self._terms = None
        return self._aff

    def __eq__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        return self._force() == o._force()  # affine, fully reduced: the representation is unique

    def __ne__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        return self._force() != o._force()

    __hash__ = None

    def to_compressed_bytes(self):
        if self._comp is None:
            lib = _rt.get_lib()
            aff = lib.upload(self._force())
            self._comp = lib.compress_aff(aff, 1)
        return self._comp

    def __str__(self):
        return self.to_compressed_bytes().hex()

    __repr__ = __str__

    @staticmethod
    def _decompress(data, check):
        data = bytes(data)
        if len(data) != 48:
            raise ValueError("serialised data seems to be invalid")
        lib = _rt.get_lib()
        aff, err = lib.decompress(data, check_subgroup=check)
        if err[0]:
            raise ValueError("serialised data seems to be invalid")
        p = G1Point._concrete(lib.download(aff, _rt.AFF))
        p._comp = data
        return p

    @staticmethod
    def from_compressed_bytes(data):
        return G1Point._decompress(data, True)

    @staticmethod
    def from_compressed_bytes_unchecked(data):
        return G1Point._decompress(data, False)

    @staticmethod
    def multiexp_unchecked(bases, scalars):
        acc = G1Point._lazy({})
        for b, s in zip(list(bases), list(scalars)):
            acc = acc + b * s
        acc._force()
        return acc

"""ORACLE (test infrastructure, never on the product path).

CPU stand-in for the ``py_arkworks_bls12381`` surface the reference imports
(stub: /root/reference/curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-54;
extra dunder names pinned by curdleproofs/curdleproofs/test_curdleproofs.py:45-128).

Two interchangeable backends hold the group arithmetic:
  * ``py``  - oracle/bls12381_py.py, Python big integers (obviously correct, slow)
  * ``c``   - oracle/cref/libbls12381_ref.so, plain C restatement (fast; used for the
              CPU baseline and for large cases), itself checked against ``py``.
Select with ``set_backend("py"|"c")`` before creating points.
"""
from . import bls12381_py as _py

R = _py.R

_BACKEND = "py"
_C = None


def set_backend(name):
    global _BACKEND, _C
    if name == "c":
        from . import cref_binding

        _C = cref_binding.load()
    elif name != "py":
        raise ValueError(name)
    _BACKEND = name


def get_backend():
    return _BACKEND


class Scalar:
    __slots__ = ("v",)

    def __init__(self, value=0):
        if isinstance(value, Scalar):
            value = value.v
        if not isinstance(value, int) or isinstance(value, bool):
            raise TypeError("Scalar() needs an int")
        if value < 0:
            raise OverflowError("can't convert negative int to unsigned")
        self.v = value % R

    @staticmethod
    def _raw(v):
        s = Scalar.__new__(Scalar)
        s.v = v
        return s

    def __add__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((self.v + o.v) % R)

    __radd__ = __add__

    def __sub__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((self.v - o.v) % R)

    def __rsub__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw((o.v - self.v) % R)

    def __mul__(self, o):
        if not isinstance(o, Scalar):
            return NotImplemented
        return Scalar._raw(self.v * o.v % R)

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
        return Scalar._raw((-self.v) % R)

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
        # inverse of zero is returned as zero: cp/util.py:51-54 relies on a value
        # coming back so that its own assert fires.
        return Scalar._raw(pow(self.v, -1, R) if self.v else 0)

    def square(self):
        return Scalar._raw(self.v * self.v % R)

    def pow(self, e):
        e = int(e) if not isinstance(e, int) else e
        return Scalar._raw(pow(self.v, e, R))

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
        if v >= R:
            raise ValueError("serialised data seems to be invalid")
        return Scalar._raw(v)


class G1Point:
    """Value is backend-specific: ``py`` -> None | (X, Y, Z); ``c`` -> 144-byte blob."""

    __slots__ = ("p", "b")

    def __init__(self):
        self.b = _BACKEND
        self.p = _py.GENERATOR if _BACKEND == "py" else _C.generator()

    @staticmethod
    def _wrap(p, b=None):
        g = G1Point.__new__(G1Point)
        g.p = p
        g.b = b or _BACKEND
        return g

    @staticmethod
    def identity():
        return G1Point._wrap(_py.INF if _BACKEND == "py" else _C.identity())

    def __add__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        if self.b == "py":
            return G1Point._wrap(_py.add(self.p, o.p), "py")
        return G1Point._wrap(_C.add(self.p, o.p), "c")

    __radd__ = __add__

    def __sub__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        if self.b == "py":
            return G1Point._wrap(_py.sub(self.p, o.p), "py")
        return G1Point._wrap(_C.sub(self.p, o.p), "c")

    def __rsub__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        return o.__sub__(self)

    def __neg__(self):
        if self.b == "py":
            return G1Point._wrap(_py.neg(self.p), "py")
        return G1Point._wrap(_C.neg(self.p), "c")

    def __mul__(self, k):
        if not isinstance(k, Scalar):
            return NotImplemented
        if self.b == "py":
            return G1Point._wrap(_py.mul(self.p, k.v), "py")
        return G1Point._wrap(_C.mul(self.p, k.v), "c")

    __rmul__ = __mul__

    def __eq__(self, o):
        if not isinstance(o, G1Point):
            return NotImplemented
        if self.b == "py":
            return _py.eq(self.p, o.p)
        return _C.eq(self.p, o.p)

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    __hash__ = None

    def to_compressed_bytes(self):
        if self.b == "py":
            return _py.compress(self.p)
        return _C.compress(self.p)

    def __str__(self):
        return bytes(self.to_compressed_bytes()).hex()

    __repr__ = __str__

    @staticmethod
    def from_compressed_bytes(data):
        if _BACKEND == "py":
            return G1Point._wrap(_py.decompress(data, check_subgroup=True))
        return G1Point._wrap(_C.decompress(bytes(data), True))

    @staticmethod
    def from_compressed_bytes_unchecked(data):
        if _BACKEND == "py":
            return G1Point._wrap(_py.decompress(data, check_subgroup=False))
        return G1Point._wrap(_C.decompress(bytes(data), False))

    @staticmethod
    def multiexp_unchecked(bases, scalars):
        bases = list(bases)
        scalars = list(scalars)
        if len(bases) != len(scalars):
            raise ValueError("bases and scalars must have the same length")
        if _BACKEND == "py":
            return G1Point._wrap(
                _py.msm_pippenger([b.p for b in bases], [s.v for s in scalars])
            )
        return G1Point._wrap(_C.msm([b.p for b in bases], [s.v for s in scalars]))

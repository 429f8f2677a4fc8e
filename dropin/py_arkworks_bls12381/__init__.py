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

Observation is BATCHED over everything that is pending.  The reference observes its points one at a time
(``transcript.append(point_projective_to_bytes(P))`` per point, cp/curdleproofs.py:100-125, cp/same_msm.py:60-70 ...),
and a lone scalar multiplication on a GPU is a latency chain of 255 dependent doublings (~2 ms whatever the batch size).
So the first observation of ANY lazy point evaluates every lazy point that is still alive (a weak registry; loop
temporaries of ``compute_MSM`` die at once and are never evaluated) in one launch sequence - scalar multiples through
cpg_g1_mul, longer combinations through cpg_g1_msm_batched per size class - and converts, compresses and downloads them
together: n = 128 ``CurdleProofsProof.new`` makes ~25 launch sequences instead of ~550.  Early evaluation cannot change a
value: leaves are immutable and a combination is a pure function of them.

Encodings met before are not decompressed again: a bounded host-side map  48-byte encoding -> affine bytes  is filled by
every decompression and every compression (cp/msm_accumulator.py:65 decodes, per verification, ~620 bases the very
process encoded a moment earlier; Whisk's pre-shuffle trackers are an earlier shuffle's post-shuffle trackers).  An
entry remembers whether the subgroup check was run, and ``from_compressed_bytes`` (checked) only trusts checked ones.

Deferred decoding (OPT-IN: ``defer_decoding(True)`` or CPG_DROPIN_DEFER_DECODE=1).  The reference decodes one point per
call (cp/whisk_interface.py:96-100: 4 ell trackers, then the proof's points), and one point on a GPU is the latency of a
single thread's 457-product square-root chain (0.46 ms): 715 such calls are 0.33 s of a 0.35 s ``.verify``.  With
deferral on, ``from_compressed_bytes[_unchecked]`` only records the 48 bytes; the first point whose VALUE is needed
decodes everything recorded so far in one launch.  The price is where a malformed encoding is reported: the same
``ValueError``, but raised by the first use of that point (arithmetic that gets observed, ``==``,
``to_compressed_bytes``) instead of by the decoding call - a point that is decoded and never used raises nothing.
``IsValidWhiskShuffleProof`` turns any exception into ``False`` (cp/whisk_interface.py:84-87), so its verdicts do not
change; the default stays eager, the wheel's exact behaviour.
"""
import os as _os
import weakref as _weakref

from curdleproofs_pie_b200 import runtime as _rt

_R = _rt.R_ORDER
_ZERO_AFF = bytes(_rt.AFF)
_INF48 = bytes([0xC0]) + bytes(47)

_pending = _weakref.WeakValueDictionary()      # id -> lazy point nobody has observed yet (G1Point is unhashable, like the wheel's)
_BATCH_MAX_TERMS = 1024            # a pending combination longer than this waits for its own observation
_BATCH_MAX_POINTS = 1 << 14
_decoded = {}                      # 48-byte encoding -> (affine bytes, subgroup-checked)
_DECODED_MAX = 1 << 17             # ~25 MB of host memory; emptied when full
_SIZE_CLASSES = (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)
_undecoded = _weakref.WeakValueDictionary()    # id -> point whose 48 bytes are recorded but not decoded yet
_defer = _os.environ.get("CPG_DROPIN_DEFER_DECODE", "") not in ("", "0")
_BAD = -1                          # G1Point._dec: None = nothing pending, 0 / 1 = decode pending (subgroup check flag), _BAD = malformed
_INVALID = "serialised data seems to be invalid"


def defer_decoding(on=True):
    """Switch deferred decoding (module docstring) on or off; returns the previous setting."""
    global _defer
    prev, _defer = _defer, bool(on)
    return prev


def _flush_decodes():
    """Decode every recorded encoding in one launch per check flag; malformed ones are marked, not raised here."""
    if not _undecoded:
        return
    lib = _rt.get_lib()
    todo = [p for p in list(_undecoded.values()) if p._aff is None and p._dec in (0, 1)]
    _undecoded.clear()
    for flag in (0, 1):
        grp = [p for p in todo if p._dec == flag]
        if not grp:
            continue
        aff, err = lib.decompress(b"".join(p._comp for p in grp), check_subgroup=bool(flag))
        raw = lib.download(aff, len(grp) * _rt.AFF)
        for i, p in enumerate(grp):
            if err[i]:
                p._dec = _BAD
            else:
                p._aff, p._dec = raw[i * _rt.AFF:(i + 1) * _rt.AFF], None
                _remember(p._comp, p._aff, bool(flag))


def _value(p):
    """Affine bytes of a concrete (possibly not yet decoded) point."""
    if p._aff is None:
        if p._dec in (0, 1):
            _undecoded[id(p)] = p                        # (still pending if an earlier flush was interrupted by an error)
            _flush_decodes()
        if p._dec == _BAD:
            raise ValueError(_INVALID)
    return p._aff


def _remember(comp, aff, checked):
    if len(_decoded) >= _DECODED_MAX:
        _decoded.clear()
    cur = _decoded.get(comp)
    if cur is None or (checked and not cur[1]):
        _decoded[comp] = (aff, checked)


class _View:
    """A window into a device buffer (a .ptr like runtime.DevBuf, no ownership)."""

    __slots__ = ("ptr",)

    def __init__(self, buf, offset):
        self.ptr = buf.ptr + offset


def _evaluate(points):
    """All of `points` (lazy, non-empty) in one launch sequence; fills _aff and _comp of each.

    Every call into the library from here is a latency chain of a few milliseconds whatever its size (a Horner pass of
    255 dependent doublings, a conversion, a compression, two downloads), so all results land in ONE Jacobian buffer that
    is converted, compressed and downloaded once.  The MSMs stay one call per size class: padding every point of a flush to
    the widest one was tried and is slower (the bucket reduction is paid per MSM and per window whatever the real number
    of terms: `.new` 0.10 - 0.26 s from run to run instead of 0.10)."""
    lib = _rt.get_lib()
    for p in points:
        for leaf, _ in p._terms.values():
            if leaf._aff is None:
                _value(leaf)                                        # decodes everything recorded; raises for a malformed leaf
    npts = len(points)
    jac_all = lib.alloc(npts * _rt.JAC)
    zero32 = bytes(32)

    def padded(grp, width):
        bl, sl = [], []
        for p in grp:
            leaves = list(p._terms.values())
            pad = width - len(leaves)
            bl.append(b"".join(leaf._aff for leaf, _ in leaves) + _ZERO_AFF * pad)
            sl.append(b"".join(k.to_bytes(32, "little") for _, k in leaves) + zero32 * pad)
        return lib.upload(b"".join(bl)), lib.upload(b"".join(sl))

    order = []                     # points in the order of their slots in jac_all
    ones = [p for p in points if len(p._terms) == 1]
    if ones:
        leaves = [next(iter(p._terms.values())) for p in ones]
        jac = lib.aff_to_jac(lib.upload(b"".join(leaf._aff for leaf, _ in leaves)), len(ones))
        sc = lib.upload(b"".join(k.to_bytes(32, "little") for _, k in leaves))
        lib.check(lib.c.cpg_g1_mul(jac.ptr, sc.ptr, len(ones), 1, jac_all.ptr), "cpg_g1_mul")
        order += ones
    by_class = {}
    for p in points:
        n = len(p._terms)
        if n > 1:
            by_class.setdefault(next((c for c in _SIZE_CLASSES if n <= c), n), []).append(p)
    for cls, grp in by_class.items():
        if cls > _SIZE_CLASSES[-1] or len(grp) == 1:         # its own length, no padding
            for p in grp:
                bases, sc = padded([p], len(p._terms))
                lib.msm_batched(bases, 0, sc, 1, len(p._terms), out=_View(jac_all, len(order) * _rt.JAC))
                order.append(p)
            continue
        w = max(len(p._terms) for p in grp)
        bases, sc = padded(grp, w)
        lib.msm_batched(bases, w, sc, len(grp), w, out=_View(jac_all, len(order) * _rt.JAC))
        order += grp
    aff = lib.jac_to_aff(jac_all, npts)
    comp = lib.compress_aff(aff, npts)
    raw = lib.download(aff, npts * _rt.AFF)
    for i, p in enumerate(order):
        p._aff = raw[i * _rt.AFF:(i + 1) * _rt.AFF]
        p._comp = comp[i * 48:(i + 1) * 48]
        p._terms = None
        _pending.pop(id(p), None)
        _remember(p._comp, p._aff, False)


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

    __slots__ = ("_aff", "_terms", "_comp", "_dec", "__weakref__")

    def __init__(self):
        lib = _rt.get_lib()
        aff = lib.jac_to_aff(lib.generator(), 1)
        self._aff = lib.download(aff, _rt.AFF)
        self._terms = None
        self._comp = None
        self._dec = None

    # -- construction helpers --
    @staticmethod
    def _concrete(aff_bytes):
        p = G1Point.__new__(G1Point)
        p._aff = aff_bytes
        p._terms = None
        p._comp = None
        p._dec = None
        return p

    @staticmethod
    def _lazy(terms):
        p = G1Point.__new__(G1Point)
        p._aff = None
        p._terms = terms
        p._comp = None
        p._dec = None
        if terms:
            _pending[id(p)] = p
        return p

    @staticmethod
    def identity():
        return G1Point._concrete(_ZERO_AFF)

    def _as_terms(self):
        """{id(leaf): (leaf, coefficient)} view of this point."""
        if self._aff is not None:
            return {} if self._aff == _ZERO_AFF else {id(self): (self, 1)}
        if self._dec is not None:                    # recorded, not decoded: never the identity (that encoding is decoded at once)
            return {id(self): (self, 1)}
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

    # -- observation: one launch sequence for everything pending --
    def _force(self):
        if self._aff is not None:
            return self._aff
        if self._dec is not None:
            return _value(self)
        if not self._terms:
            self._aff, self._comp, self._terms = _ZERO_AFF, _INF48, None
            return self._aff
        batch = [self]
        for p in list(_pending.values()):
            if p is not self and p._aff is None and p._dec is None and p._terms and len(p._terms) <= _BATCH_MAX_TERMS and len(batch) < _BATCH_MAX_POINTS:
                batch.append(p)
        _evaluate(batch)
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
        if self._comp is None or self._dec is not None:
            self._force()
        if self._comp is None:                       # a concrete point that was never encoded (the generator)
            lib = _rt.get_lib()
            self._comp = lib.compress_aff(lib.upload(self._aff), 1)
            _remember(self._comp, self._aff, False)
        return self._comp

    def __str__(self):
        return self.to_compressed_bytes().hex()

    __repr__ = __str__

    @staticmethod
    def _decompress(data, check):
        data = bytes(data)
        if len(data) != 48:
            raise ValueError(_INVALID)
        known = _decoded.get(data)
        if known is not None and (known[1] or not check):
            p = G1Point._concrete(known[0])
            p._comp = data
            return p
        if _defer and not (data[0] & 0x40):          # infinity encodings (and their malformed variants) are settled at once
            p = G1Point._concrete(None)
            p._comp, p._dec = data, 1 if check else 0
            _undecoded[id(p)] = p
            return p
        lib = _rt.get_lib()
        aff, err = lib.decompress(data, check_subgroup=check)
        if err[0]:
            raise ValueError("serialised data seems to be invalid")
        p = G1Point._concrete(lib.download(aff, _rt.AFF))
        p._comp = data
        _remember(data, p._aff, bool(check))
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

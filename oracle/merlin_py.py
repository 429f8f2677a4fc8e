"""ORACLE (test infrastructure, never on the product path).

Restatement of the reference's Fiat-Shamir transcript stack:
  Keccak-f[1600]            merlin_transcripts/merlin_transcripts/keccak.py:56-66
  STROBE-128 (R = 166)      merlin_transcripts/merlin_transcripts/strobe.py:16-107
  Merlin v1.0 framing       merlin_transcripts/merlin_transcripts/merlin_transcript.py:6-24
  scalar challenges         curdleproofs/curdleproofs/curdleproofs_transcript.py:7-28
Pinned by the reference's KATs (merlin_transcripts/merlin_transcripts/test_merlin.py:18,29,40)
in tests/test_oracle_kat.py.
"""
from .bls12381_py import R as _R

_MASK = (1 << 64) - 1
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
# rotation offsets indexed x + 5*y
_ROT = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]

_keccak_native = None


def use_native_keccak(fn):
    """Swap in a C permutation (oracle/cref) to make long oracle runs bearable."""
    global _keccak_native
    _keccak_native = fn


def keccak_f1600(state):
    if _keccak_native is not None:
        return _keccak_native(state)
    a = [int.from_bytes(state[8 * i:8 * i + 8], "little") for i in range(25)]
    for rc in _RC:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ (((c[(x + 1) % 5] << 1) | (c[(x + 1) % 5] >> 63)) & _MASK) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                v = a[x + 5 * y]
                r = _ROT[x + 5 * y]
                b[y + 5 * ((2 * x + 3 * y) % 5)] = ((v << r) | (v >> (64 - r))) & _MASK if r else v
        a = [b[i] ^ (~b[(i % 5 + 1) % 5 + 5 * (i // 5)] & _MASK & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= rc
    out = bytearray(200)
    for i in range(25):
        out[8 * i:8 * i + 8] = a[i].to_bytes(8, "little")
    return out


RATE = 166
F_I, F_A, F_C, F_T, F_M, F_K = 1, 2, 4, 8, 16, 32


class Strobe:
    def __init__(self, protocol_label):
        st = bytearray(200)
        st[0:6] = bytes([1, RATE + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.st = keccak_f1600(st)
        self.pos = 0
        self.pos_begin = 0
        self.flags = 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[RATE + 1] ^= 0x80
        self.st = keccak_f1600(self.st)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.st[self.pos] ^= byte
            self.pos += 1
            if self.pos == RATE:
                self._run_f()

    def _begin(self, flags, more):
        if more:
            assert self.flags == flags
            return
        assert not flags & F_T
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (F_C | F_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin(F_M | F_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin(F_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin(F_I | F_A | F_C, more)
        out = bytearray(n)
        for i in range(n):
            out[i] = self.st[self.pos]
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == RATE:
                self._run_f()
        return out

    def key(self, data, more):
        self._begin(F_A | F_C, more)
        for byte in data:
            self.st[self.pos] = byte
            self.pos += 1
            if self.pos == RATE:
                self._run_f()


class Transcript:
    """Merlin framing + the curdleproofs scalar-challenge rule."""

    def __init__(self, label):
        self.s = Strobe(b"Merlin v1.0")
        self.append(b"dom-sep", label)

    def append(self, label, msg):
        self.s.meta_ad(label, False)
        self.s.meta_ad(len(msg).to_bytes(4, "little"), True)
        self.s.ad(msg, False)

    def append_all(self, label, msgs):
        for m in msgs:
            self.append(label, m)

    def challenge_bytes(self, label, n):
        self.s.meta_ad(label, False)
        self.s.meta_ad(n.to_bytes(4, "little"), True)
        return bytes(self.s.prf(n, False))

    def challenge_int(self, label):
        """Rejection-sample a non-zero scalar < r, then bind it back into the transcript
        (curdleproofs_transcript.py:15-25)."""
        while True:
            raw = self.challenge_bytes(label, 32)
            v = int.from_bytes(raw, "little")
            if v >= _R or v == 0:
                continue
            self.append(label, raw)
            return v

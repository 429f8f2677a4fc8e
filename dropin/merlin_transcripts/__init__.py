"""Drop-in ``merlin_transcripts`` backed by libcpg.so's own STROBE-128 / Keccak-f[1600] (csrc/host_transcript.h) -
SURVEY 8 f-1 for callers that run the UNMODIFIED reference: its pure-Python Keccak costs ~0.7 ms per permutation,
i.e. ~0.5 s of Fiat-Shamir per n = 128 prove or verify; here every append / challenge is one ctypes call.

Surface: /root/reference/merlin_transcripts/merlin_transcripts/merlin_transcript.py:6-24 (``MerlinTranscript(label)``,
``append_message``, ``append_u64``, ``challenge_bytes``); the reference's ``CurdleproofsTranscript``
(curdleproofs/curdleproofs/curdleproofs_transcript.py:7-28) subclasses it unchanged.  Put ``dropin/`` ahead of the
reference's own package on ``sys.path``.  The transcript handle lives in host memory inside the library
(cpg_merlin_new/append/challenge, include/cpg.h); there is no Python fallback.
"""
import ctypes as _c

from curdleproofs_pie_b200 import runtime as _rt

__all__ = ["MerlinTranscript"]


class MerlinTranscript:
    __slots__ = ("_h", "_lib")

    def __init__(self, label):
        self._lib = _rt.get_lib()
        label = bytes(label)
        self._h = self._lib.c.cpg_merlin_new(label, len(label))
        if not self._h:
            raise _rt.CpgError("cpg_merlin_new failed: " + self._lib.last_error())

    def append_message(self, label, message):
        label, message = bytes(label), bytes(message)
        self._lib.check(self._lib.c.cpg_merlin_append(self._h, label, len(label), message, len(message)), "cpg_merlin_append")

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, length):
        label = bytes(label)
        out = _c.create_string_buffer(max(1, int(length)))
        self._lib.check(self._lib.c.cpg_merlin_challenge(self._h, label, len(label), out, int(length)), "cpg_merlin_challenge")
        return out.raw[:length]

    def __copy__(self):
        t = object.__new__(type(self))
        t._lib = self._lib
        t._h = self._lib.c.cpg_merlin_clone(self._h)
        return t

    def __deepcopy__(self, memo):
        return self.__copy__()

    def __del__(self):
        try:
            if self._h:
                self._lib.c.cpg_merlin_free(self._h)
                self._h = None
        except Exception:
            pass

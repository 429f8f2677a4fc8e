"""The PRODUCT's STROBE-128 / Merlin / Keccak (csrc/host_transcript.h - the code the batched prover and verifier run
per proof) against the reference's own known-answer vectors, merlin_transcripts/test_merlin.py:18,29,40, and against
the oracle's Python restatement on random scripts.  CPU tier: on the host and in the host-emulated kernel; the GPU
tier runs the same cases in a real one-thread kernel (test_gpu_parity.py::test_transcript_kats_on_device)."""
import random

import pytest

from oracle import merlin_py

STROBE_INIT, META_AD, AD, PRF, KEY, MERLIN_INIT, APPEND, CHALLENGE = range(8)


def strobe_conformance(lib, on_device):
    """mt/test_merlin.py:18-37 (the STROBE conformance vector, incl. the KEY operation)"""
    out = lib.merlin_script([
        (STROBE_INIT, 0, b"", b"Conformance Test Protocol"),
        (META_AD, 0, b"", b"ms"), (META_AD, 1, b"", b"g"), (AD, 0, b"", bytes([99]) * 1024),
        (META_AD, 0, b"", b"prf"), (PRF, 0, b"", 32),
    ], on_device)
    assert out.hex() == "b48e645ca17c667fd5206ba57a6a228d72d8e1903814d3f17f622996d7cfefb0"
    out2 = lib.merlin_script([
        (STROBE_INIT, 0, b"", b"Conformance Test Protocol"),
        (META_AD, 0, b"", b"ms"), (META_AD, 1, b"", b"g"), (AD, 0, b"", bytes([99]) * 1024),
        (META_AD, 0, b"", b"prf"), (PRF, 0, b"", 32),
        (META_AD, 0, b"", b"key"), (KEY, 0, b"", out),
        (META_AD, 0, b"", b"prf"), (PRF, 0, b"", 32),
    ], on_device)
    assert out2[32:].hex() == "07e45cce8078cee259e3e375bb85d75610e2d1e1201c5f645045a194edd49ff8"


def merlin_simple(lib, on_device):
    """mt/test_merlin.py:40-47 (equivalence_simple of the Rust merlin crate)"""
    out = lib.merlin_script([(MERLIN_INIT, 0, b"", b"test protocol"), (APPEND, 0, b"some label", b"some data"), (CHALLENGE, 0, b"challenge", 32)], on_device)
    assert out.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def merlin_random_scripts(lib, on_device, seed, rounds=12):
    """long interleavings of appends and challenges (message sizes around the 166-byte rate) against the oracle"""
    rng = random.Random(seed)
    t = merlin_py.Transcript(b"curdleproofs")
    recs = [(MERLIN_INIT, 0, b"", b"curdleproofs")]
    want = b""
    for _ in range(rounds):
        for _ in range(rng.randrange(1, 6)):
            label = bytes(rng.randrange(97, 123) for _ in range(rng.randrange(1, 24)))
            msg = rng.randbytes(rng.choice([0, 1, 32, 48, 165, 166, 167, 333, 1000]))
            t.append(label, msg)
            recs.append((APPEND, 0, label, msg))
        n = rng.choice([1, 32, 64, 166, 200])
        want += t.challenge_bytes(b"chal", n)
        recs.append((CHALLENGE, 0, b"chal", n))
    assert lib.merlin_script(recs, on_device) == want


@pytest.mark.parametrize("on_device", [False, True])
def test_reference_kats(seam_lib, on_device):
    strobe_conformance(seam_lib, on_device)
    merlin_simple(seam_lib, on_device)


@pytest.mark.parametrize("on_device", [False, True])
def test_random_scripts_match_oracle(seam_lib, on_device):
    for seed in range(4):
        merlin_random_scripts(seam_lib, on_device, seed)


def test_stateful_handle_matches_script(seam_lib):
    import ctypes

    c = seam_lib.c
    h = c.cpg_merlin_new(b"test protocol", 13)
    c.cpg_merlin_append(h, b"some label", 10, b"some data", 9)
    h2 = c.cpg_merlin_clone(h)
    out = ctypes.create_string_buffer(32)
    c.cpg_merlin_challenge(h, b"challenge", 9, out, 32)
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    c.cpg_merlin_challenge(h2, b"challenge", 9, out, 32)                     # the clone is an independent fork
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    c.cpg_merlin_free(h); c.cpg_merlin_free(h2)


def test_malformed_script_is_an_error(seam_lib):
    from curdleproofs_pie_b200 import runtime as rt

    with pytest.raises(rt.CpgError):
        seam_lib.merlin_script([(CHALLENGE, 0, b"x", 64)], False, cap=8)     # output does not fit
    import ctypes

    out = ctypes.create_string_buffer(8); got = ctypes.c_size_t()
    assert seam_lib.c.cpg_merlin_script(b"\x06\x00\xff\xff", 4, 0, out, 8, ctypes.byref(got)) != 0   # truncated record

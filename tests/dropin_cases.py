"""Cases for the reference-facing ``py_arkworks_bls12381`` drop-in (dropin/), shared by the CPU tier
(host-emulated kernels) and the GPU tier.  They mirror the reference's own surface tests
(cp/test_curdleproofs.py:132-213, :233-236)."""
import random

import pytest

GEN_HEX = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
G99_HEX = "aa10e1055b14a89cc3261699524998732fddc4f30c76c1057eb83732a01416643eb015a932e4080c86f42e485973d240"
G2_HEX = "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e"   # 2 G, 3 G: Ethereum consensus BLS
G3_HEX = "89ece308f9d1f0131765212deca99697b112d61f9be9a5f1f3780a51335b3ff981747a0b2ca2179b96d2c0c9024e5224"   # vectors (tests/test_oracle_kat.py)
CURVE_ORDER = 52435875175126190479447740508185965837690552500527637822603658699938581184513

REQUIRED_G1 = {"__add__", "__sub__", "__neg__", "__mul__", "__eq__", "__ne__", "__radd__", "__rmul__", "__rsub__", "__str__",
               "identity", "to_compressed_bytes", "from_compressed_bytes", "from_compressed_bytes_unchecked", "multiexp_unchecked"}
REQUIRED_FR = {"__add__", "__sub__", "__neg__", "__mul__", "__eq__", "__ne__", "__radd__", "__rmul__", "__rsub__", "__int__", "__truediv__",
               "__rtruediv__", "from_le_bytes", "to_le_bytes", "inverse", "is_zero", "square", "pow"}


def surface_kats(mod):
    G1Point, Scalar = mod.G1Point, mod.Scalar
    assert REQUIRED_G1 <= set(dir(G1Point)) and REQUIRED_FR <= set(dir(Scalar))
    gen, ident = G1Point(), G1Point.identity()
    assert gen == gen and gen != ident
    assert (gen + gen) - gen == gen
    assert -gen + gen == ident
    assert gen * Scalar(4) == gen + gen + gen + gen
    cb = gen.to_compressed_bytes()
    assert G1Point.from_compressed_bytes(cb) == G1Point.from_compressed_bytes_unchecked(cb) == gen
    assert str(gen) == GEN_HEX
    assert bytes((gen * Scalar(99)).to_compressed_bytes()).hex() == G99_HEX
    assert bytes((gen + gen).to_compressed_bytes()).hex() == G2_HEX and bytes((gen * Scalar(3)).to_compressed_bytes()).hex() == G3_HEX
    assert G1Point.from_compressed_bytes(bytes.fromhex(G3_HEX)) - G1Point.from_compressed_bytes(bytes.fromhex(G2_HEX)) == gen
    assert bytes(ident.to_compressed_bytes()) == bytes([0xC0]) + bytes(47)
    assert (gen * Scalar(0)) == ident and (ident * Scalar(5)) == ident
    with pytest.raises(TypeError):
        {gen: True}
    with pytest.raises(ValueError):
        G1Point.from_compressed_bytes_unchecked(bytes(48))
    with pytest.raises(ValueError):
        G1Point.from_compressed_bytes_unchecked(bytes(47))
    assert bytes(Scalar(4).to_le_bytes()) == bytes.fromhex("04" + "00" * 31)
    assert int(Scalar(CURVE_ORDER - 1)) == CURVE_ORDER - 1 and int(Scalar(CURVE_ORDER)) == 0
    assert int(Scalar(2**257)) == 2**257 % CURVE_ORDER
    with pytest.raises(ValueError):
        Scalar.from_le_bytes(CURVE_ORDER.to_bytes(32, "little"))
    assert Scalar(7) * Scalar(7).inverse() == Scalar(1)
    assert Scalar(0).inverse() * Scalar(0) != Scalar(1)


def multiexp_matches_oracle(mod, cref, n, seed=11):
    from oracle import bls12381_py as bp

    rng = random.Random(seed)
    G1Point, Scalar = mod.G1Point, mod.Scalar
    ks = [rng.randrange(bp.R) for _ in range(n)]
    ss = [rng.randrange(bp.R) for _ in range(n)]
    blobs = cref.mul_batch([cref.generator()] * n, ks)
    enc = cref.compress_batch(blobs)
    pts = [G1Point.from_compressed_bytes_unchecked(e) for e in enc]
    got = G1Point.multiexp_unchecked(pts, [Scalar(s) for s in ss])
    assert bytes(got.to_compressed_bytes()) == cref.compress(cref.msm(blobs, ss))
    # the reference's compute_MSM loop (cp/msm_accumulator.py:6-12) gives the same bytes
    acc = G1Point.identity()
    for p, s in zip(pts, ss):
        acc = acc + p * Scalar(s)
    assert acc == got


def deferred_decoding(mod, cref, n=40, seed=21):
    """Opt-in deferred decoding (dropin docstring): decoding calls only record the bytes, the first use decodes everything
    recorded in one launch; values, bytes and equality are those of the eager mode, and a malformed encoding raises the
    same ValueError - at its first use instead of at the decoding call."""
    from oracle import bls12381_py as bp

    rng = random.Random(seed)
    G1Point, Scalar = mod.G1Point, mod.Scalar
    ks = [rng.randrange(1, bp.R) for _ in range(n)]
    ss = [rng.randrange(bp.R) for _ in range(n)]
    blobs = cref.mul_batch([cref.generator()] * n, ks)
    enc = cref.compress_batch(blobs)
    want = cref.compress(cref.msm(blobs, ss))
    x_off = next(x for x in range(1, 200) if pow((x**3 + 4) % bp.P, (bp.P - 1) // 2, bp.P) != 1)     # x^3 + 4 a non-residue
    off_curve = bytes([0x80]) + x_off.to_bytes(48, "big")[1:]
    getattr(mod, "_decoded", {}).clear()
    prev = mod.defer_decoding(True)
    try:
        launches0 = mod._rt.get_lib().launch_count()
        pts = [G1Point.from_compressed_bytes_unchecked(e) for e in enc]
        chk = [G1Point.from_compressed_bytes(e) for e in enc[:5]]
        assert mod._rt.get_lib().launch_count() == launches0                 # nothing ran yet
        assert all(p._aff is None for p in pts)
        got = G1Point.multiexp_unchecked(pts, [Scalar(s) for s in ss])
        assert bytes(got.to_compressed_bytes()) == want
        assert all(p._aff is not None for p in pts + chk)                     # one flush decoded every recorded point
        assert [bytes(p.to_compressed_bytes()) for p in pts] == [bytes(e) for e in enc]
        assert chk[0] == pts[0] and chk[1] != pts[0]
        # identity encodings and wrong lengths are settled at once, as in eager mode
        assert G1Point.from_compressed_bytes_unchecked(bytes([0xC0]) + bytes(47)) == G1Point.identity()
        with pytest.raises(ValueError):
            G1Point.from_compressed_bytes_unchecked(bytes([0xC0]) + bytes(46) + b"\x01")
        with pytest.raises(ValueError):
            G1Point.from_compressed_bytes_unchecked(bytes(47))
        # malformed encodings: recorded silently, reported by the first use - and only by uses of THAT point
        getattr(mod, "_decoded", {}).clear()
        bad = G1Point.from_compressed_bytes_unchecked(off_curve)
        good = G1Point.from_compressed_bytes_unchecked(enc[0])
        assert bytes((good * Scalar(3)).to_compressed_bytes()) == cref.compress(cref.mul_batch([blobs[0]], [3])[0])
        for use in (lambda: bad.to_compressed_bytes(), lambda: bad == good, lambda: (bad * Scalar(2) + good).to_compressed_bytes()):
            with pytest.raises(ValueError):
                use()
        # a point off the r-order subgroup passes the unchecked decoder and fails the checked one, deferred or not
        h = next(bytes([0x80]) + x.to_bytes(48, "big")[1:] for x in range(1, 400)
                 if pow((x**3 + 4) % bp.P, (bp.P - 1) // 2, bp.P) == 1 and x != x_off)
        G1Point.from_compressed_bytes_unchecked(h).to_compressed_bytes()
        with pytest.raises(ValueError):
            G1Point.from_compressed_bytes(h).to_compressed_bytes()
    finally:
        mod.defer_decoding(prev)
    # eager again: the same malformed encoding raises in the decoding call
    getattr(mod, "_decoded", {}).clear()
    with pytest.raises(ValueError):
        G1Point.from_compressed_bytes_unchecked(off_curve)

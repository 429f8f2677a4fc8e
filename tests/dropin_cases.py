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

"""The device field arithmetic (csrc/bigint.cuh, field.cuh) compiled for the host with the PTX carry
flag emulated, against Python integers: Montgomery product, dedicated squaring, add/sub, inversion,
square root, for Fq (12 limbs) and Fr (8 limbs), with limb patterns chosen to stress carries."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@pytest.fixture(scope="module")
def hs():
    out = os.path.join(ROOT, "tests", "_build", "libhostfield.so")
    src = os.path.join(ROOT, "tests", "host_seam", "host_seam.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-x", "c++", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


def call(lib, name, nbytes, *vals):
    o = ctypes.create_string_buffer(nbytes)
    getattr(lib, name)(*[v.to_bytes(nbytes, "little") for v in vals], o)
    return int.from_bytes(o.raw, "little")


def patterns(mod, nlimbs, rng, count):
    ones = (1 << 32) - 1
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, 1 << (32 * nlimbs - 4)]
    for _ in range(count):
        kind = rng.randrange(4)
        if kind == 0:
            v = rng.randrange(mod)
        elif kind == 1:        # limbs drawn from {0, 1, 0x7fffffff, 0x80000000, 0xffffffff}
            v = sum(rng.choice([0, 1, 0x7FFFFFFF, 0x80000000, ones]) << (32 * i) for i in range(nlimbs)) % mod
        elif kind == 2:
            v = rng.randrange(mod) >> rng.randrange(32 * nlimbs)
        else:
            v = (mod - 1 - rng.randrange(1 << 40)) % mod
        vals.append(v)
    return vals


@pytest.mark.parametrize("field,mod,nlimbs", [("fq", P, 12), ("fr", R, 8)])
def test_mul_sqr_add_sub(hs, field, mod, nlimbs):
    rng = random.Random(nlimbs)
    nb = 4 * nlimbs
    Ri = pow(pow(2, 32 * nlimbs, mod), -1, mod)
    vals = patterns(mod, nlimbs, rng, 6000)
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert call(hs, "hs_%s_mul" % field, nb, a, b) == a * b * Ri % mod
        assert call(hs, "hs_%s_sqr" % field, nb, a) == a * a * Ri % mod
        assert call(hs, "hs_%s_add" % field, nb, a, b) == (a + b) % mod
        assert call(hs, "hs_%s_sub" % field, nb, a, b) == (a - b) % mod


def test_inverse_sqrt_montgomery_form(hs):
    rng = random.Random(5)
    Rq = pow(2, 384, P)
    Rr = pow(2, 256, R)
    for _ in range(20):
        a = rng.randrange(1, P)
        am = a * Rq % P
        assert call(hs, "hs_fq_to_mont", 48, a) == am and call(hs, "hs_fq_from_mont", 48, am) == a
        assert call(hs, "hs_fq_inv", 48, am) == pow(a, -1, P) * Rq % P
        sq = a * a % P
        y = call(hs, "hs_fq_sqrt", 48, sq * Rq % P) * pow(Rq, -1, P) % P
        assert y * y % P == sq
        k = rng.randrange(1, R)
        assert call(hs, "hs_fr_inv", 32, k * Rr % R) == pow(k, -1, R) * Rr % R


def test_lazy_chain_forms(hs):
    """mont_mul_n / mont_sqr_n with LAZY = true (the square-root chain of Decompress): operands anywhere in [0, 2p),
    no final subtraction, result congruent to a b / R and below 2p again (in fact < 1.41 p); reduce_once ends a chain.
    The whole sliding-window exponentiation on top of them equals Python's pow for residues, non-residues and 0."""
    rng = random.Random(77)
    Ri = pow(pow(2, 384, P), -1, P)
    vals = patterns(2 * P, 12, rng, 6000) + [P, P + 1, P - 1, 2 * P - 1, 2 * P - 2]
    for i, a in enumerate(vals):
        b = vals[(i * 11 + 5) % len(vals)]
        m = call(hs, "hs_fq_mul_lazy", 48, a, b)
        assert m < 2 * P and m % P == a * b * Ri % P
        q = call(hs, "hs_fq_sqr_lazy", 48, a)
        assert q < 2 * P and q % P == a * a * Ri % P
        assert call(hs, "hs_fq_reduce_once", 48, a) == a % P
    Rq = pow(2, 384, P)
    for a in patterns(P, 12, rng, 300):
        want = pow(a * pow(Rq, -1, P) % P, (P + 1) // 4, P) * Rq % P          # a is a Montgomery form
        assert call(hs, "hs_fq_sqrt", 48, a) == want


def test_safegcd_inversion_matches_fermat_and_python(hs):
    """fq_inv is the Bernstein-Yang division-step inversion on signed 30-bit limbs (field.cuh); cross-checked against
    the Fermat exponentiation it replaced and against Python on carry-stressing limb patterns, 0 and the extremes"""
    rng = random.Random(381)
    Rq = pow(2, 384, P)
    vals = patterns(P, 12, rng, 3000) + [P - 1, 1, 2, 3, (1 << 380), (1 << 381) % P, Rq, pow(Rq, -1, P)]
    for am in vals:
        got = call(hs, "hs_fq_inv", 48, am)
        if am == 0:
            assert got == 0
            continue
        a = am * pow(Rq, -1, P) % P                       # am is the Montgomery form of a
        assert got == pow(a, -1, P) * Rq % P, hex(am)
    for am in vals[:200]:
        assert call(hs, "hs_fq_inv", 48, am) == call(hs, "hs_fq_inv_fermat", 48, am)

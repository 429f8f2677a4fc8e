"""ORACLE (test infrastructure, never on the product path).

Pure-Python big-integer restatement of the BLS12-381 G1 / Fr arithmetic that the
reference delegates to the external wheel ``py_arkworks_bls12381`` 0.3.5
(pinned at /root/reference/curdleproofs/pyproject.toml:10 and
curdleproofs/poetry.lock:233-234; its Rust source is NOT under /root/reference,
only the stub curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-54).

What is restated is the published mathematics, not a source file:
  * curve  y^2 = x^3 + 4 over Fq, prime-order subgroup of order r
  * ZCash/IETF compressed G1 encoding (48 B big-endian x, flag bits 0x80/0x40/0x20)
  * Fr canonical 32-byte little-endian encoding

Pinned against the reference's own known-answer tests
(curdleproofs/curdleproofs/test_curdleproofs.py:179-180, :196, :201-213, :236)
in tests/test_oracle_kat.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B_COEFF = 4
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
HALF_P = (P - 1) // 2
SQRT_EXP = (P + 1) // 4

# A point is None (identity) or a Jacobian triple (X, Y, Z) with Z != 0.
INF = None


def is_inf(pt):
    return pt is None


def from_affine(x, y):
    return (x % P, y % P, 1)


def to_affine(pt):
    if pt is None:
        return None
    X, Y, Z = pt
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def neg(pt):
    if pt is None:
        return None
    X, Y, Z = pt
    return (X, (-Y) % P, Z)


def double(pt):
    if pt is None:
        return None
    X, Y, Z = pt
    if Y == 0:
        return None
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return double(p1)
        return None
    H = (U2 - U1) % P
    I = (2 * H) * (2 * H) % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def sub(p1, p2):
    return add(p1, neg(p2))


def eq(p1, p2):
    if p1 is None or p2 is None:
        return p1 is None and p2 is None
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    if X1 * Z2Z2 % P != X2 * Z1Z1 % P:
        return False
    return Y1 * Z2 * Z2Z2 % P == Y2 * Z1 * Z1Z1 % P


def mul(pt, k):
    """k * pt, MSB-first double-and-add (what G1Projective * Fr does)."""
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = double(acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc


def is_on_curve(x, y):
    return (y * y - x * x * x - B_COEFF) % P == 0


def in_subgroup(pt):
    # full-order check: r * P == identity (mul() reduces mod r, so do it by hand)
    acc = None
    for bit in bin(R)[2:]:
        acc = double(acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc is None


def compress(pt):
    """48-byte ZCash encoding (reference call site cp/util.py:27-28)."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = to_affine(pt)
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > HALF_P:
        out[0] |= 0x20
    return bytes(out)


def decompress(data, check_subgroup=False):
    """Inverse of compress; raises ValueError on any malformed encoding
    (reference call sites cp/util.py:35-36, cp/msm_accumulator.py:65)."""
    data = bytes(data)
    if len(data) != 48:
        raise ValueError("compressed G1 point must be 48 bytes")
    flags = data[0]
    if not flags & 0x80:
        raise ValueError("compression flag not set")
    x = int.from_bytes(data, "big") & ((1 << 381) - 1)
    if flags & 0x40:
        if x != 0 or flags & 0x20:
            raise ValueError("non-canonical infinity encoding")
        return None
    if x >= P:
        raise ValueError("x coordinate not in field")
    rhs = (x * x * x + B_COEFF) % P
    y = pow(rhs, SQRT_EXP, P)
    if y * y % P != rhs:
        raise ValueError("x is not on the curve")
    if (y > HALF_P) != bool(flags & 0x20):
        y = P - y
    pt = (x, y, 1)
    if check_subgroup and not in_subgroup(pt):
        raise ValueError("point not in the prime-order subgroup")
    return pt


GENERATOR = (GX, GY, 1)


def msm_naive(points, scalars):
    """The reference's compute_MSM loop (cp/msm_accumulator.py:6-12)."""
    acc = None
    for pt, k in zip(points, scalars):
        acc = add(acc, mul(pt, k))
    return acc


def msm_pippenger(points, scalars, c=None):
    """Bucket method, unsigned digits; independent second route to the same value."""
    pts = list(points)
    ks = [k % R for k in scalars]
    n = min(len(pts), len(ks))
    if n == 0:
        return None
    if c is None:
        c = 3 if n < 32 else max(3, (n.bit_length() * 69) // 100 + 2)
    total = None
    nwin = (255 + c - 1) // c
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = double(total)
        buckets = [None] * (1 << c)
        for pt, k in zip(pts[:n], ks[:n]):
            d = (k >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = add(buckets[d], pt)
        run = None
        acc = None
        for d in range((1 << c) - 1, 0, -1):
            run = add(run, buckets[d])
            acc = add(acc, run)
        total = add(total, acc)
    return total

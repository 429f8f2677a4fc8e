"""Parity cases shared by the CPU tier (host-emulated kernels, tests/test_seam_parity.py) and the
GPU tier (tests/test_gpu_parity.py).  Each case drives the C ABI (include/cpg.h) through
curdleproofs_pie_b200.runtime and compares compressed bytes with the CPU oracle (oracle/).
Bit-exact: integer/byte work, no tolerance."""
import random

from curdleproofs_pie_b200 import runtime as rt
from oracle import bls12381_py as bp

R = bp.R
INF48 = bytes([0xC0]) + bytes(47)


def rand_points(cref, rng, k, with_identity=False):
    """k oracle points (144-B oracle blobs) and their 48-B encodings."""
    g = cref.generator()
    ks = [rng.randrange(1, R) for _ in range(k)]
    blobs = cref.mul_batch([g] * k, ks)
    if with_identity and k >= 3:
        blobs[k // 2] = cref.identity()
    return blobs, cref.compress_batch(blobs)


def upload_points(lib, enc):
    """compressed encodings -> (affine DevBuf, jacobian DevBuf) through the device decompressor."""
    aff, err = lib.decompress(b"".join(enc))
    assert not any(err), err
    return aff, lib.aff_to_jac(aff, len(enc))


def split48(data):
    return [data[i:i + 48] for i in range(0, len(data), 48)]


def edge_scalars(rng, k):
    base = [0, 1, 2, R - 1, R - 2, (1 << 255) % R, (1 << 128), (1 << 128) - 1, 0x8888888888888888888888888888888888888888888888888888888888888888 % R]
    return (base + [rng.randrange(R) for _ in range(max(0, k - len(base)))])[:k]


def case_empty(lib, crs_hex, ell):
    """Zero-length calls through every batch entry point: no launch is needed, nothing is written, nothing fails
    (the reference's list comprehensions over empty sequences, cp/whisk_interface.py:96-100, simply produce [])."""
    from curdleproofs_pie_b200 import whisk

    buf, _ = lib.decompress(b"")
    assert lib.decompress(b"")[1] == []
    assert lib.compress_jac(lib.alloc(144), 0) == b"" and lib.compress_aff(lib.alloc(96), 0) == b""
    lib.mul(lib.alloc(144), lib.alloc(32), 0)
    lib.jac_to_aff(lib.alloc(144), 0)
    lib.aff_to_jac(lib.alloc(96), 0)
    lib.add(lib.alloc(144), lib.alloc(144), 0)
    assert lib.eq(lib.alloc(144), lib.alloc(144), 0) == [] and lib.is_identity(lib.alloc(144), 0) == []
    crs = bytes.fromhex(crs_hex)
    ver = whisk.BatchVerifier(crs, ell, lib=lib)
    assert ver.verify([], []) == [] and ver.verify_raw(b"", b"", 0) == b""
    ver.close()
    prover = whisk.BatchProver(crs, ell, lib=lib)
    assert list(prover.prove_drawn([], random.Random(1))) == []
    prover.close()
    lib.sync()


def case_roundtrip(lib, cref, k, seed=1):
    rng = random.Random(seed)
    blobs, enc = rand_points(cref, rng, k, with_identity=True)
    aff, jac = upload_points(lib, enc)
    assert split48(lib.compress_aff(aff, k)) == enc
    assert split48(lib.compress_jac(jac, k)) == enc
    # checked variant accepts subgroup points
    aff2, err = lib.decompress(b"".join(enc[:4]), check_subgroup=True)
    assert err == [0, 0, 0, 0]
    # malformed encodings are reported per element, neighbours unaffected
    x = 1
    while pow(x * x * x + 4, (bp.P - 1) // 2, bp.P) == 1:
        x += 1
    off_curve = bytearray(x.to_bytes(48, "big")); off_curve[0] |= 0x80
    bad = [bytes([enc[0][0] & 0x7F]) + enc[0][1:], bytes([0xC0]) + bytes(46) + b"\x01", bytes([0xE0]) + bytes(47),
           bytes([0x9F]) + b"\xff" * 47, bytes(off_curve), enc[1]]
    _, err = lib.decompress(b"".join(bad))
    assert err == [1, 2, 2, 3, 4, 0]
    # a curve point outside the r-order subgroup: rejected only by the checked variant
    xx = 2
    while True:
        rhs = (xx ** 3 + 4) % bp.P
        y = pow(rhs, (bp.P + 1) // 4, bp.P)
        if y * y % bp.P == rhs and not bp.in_subgroup((xx, y, 1)):
            break
        xx += 1
    e = bp.compress((xx, y, 1))
    assert lib.decompress(e, check_subgroup=False)[1] == [0]
    assert lib.decompress(e, check_subgroup=True)[1] == [5]


def case_group_law(lib, cref, k, seed=2):
    rng = random.Random(seed)
    a_blob, a_enc = rand_points(cref, rng, k, with_identity=True)
    b_blob, b_enc = rand_points(cref, rng, k)
    # edge lanes: P + P (doubling branch), P + (-P), P + identity
    b_blob[0], b_enc[0] = a_blob[0], a_enc[0]
    b_blob[1] = cref.neg(a_blob[1]); b_enc[1] = cref.compress(b_blob[1])
    b_blob[2] = cref.identity(); b_enc[2] = INF48
    _, ja = upload_points(lib, a_enc)
    _, jb = upload_points(lib, b_enc)
    want_add = [cref.compress(cref.add(x, y)) for x, y in zip(a_blob, b_blob)]
    want_sub = [cref.compress(cref.sub(x, y)) for x, y in zip(a_blob, b_blob)]
    assert split48(lib.compress_jac(lib.add(ja, jb, k), k)) == want_add
    assert split48(lib.compress_jac(lib.sub(ja, jb, k), k)) == want_sub
    assert split48(lib.compress_jac(lib.neg(ja, k), k)) == [cref.compress(cref.neg(x)) for x in a_blob]
    assert lib.eq(ja, jb, k) == [1 if cref.eq(x, y) else 0 for x, y in zip(a_blob, b_blob)]
    # projectively different representations of equal points compare equal
    s = lib.add(lib.sub(ja, jb, k), jb, k)
    assert lib.eq(s, ja, k) == [1] * k
    assert lib.is_identity(lib.sub(ja, ja, k), k) == [1] * k
    ks = edge_scalars(rng, k)
    dk = lib.upload(rt.scalars_to_bytes(ks))
    got = split48(lib.compress_jac(lib.mul(ja, dk, k), k))
    assert got == [cref.compress(cref.mul(x, s)) for x, s in zip(a_blob, ks)]


def case_fold(lib, cref, rows, m, seed=3):
    rng = random.Random(seed)
    L_blob, L_enc = rand_points(cref, rng, rows * m)
    R_blob, R_enc = rand_points(cref, rng, rows * m, with_identity=True)
    xs = [rng.randrange(R) for _ in range(rows)]
    _, jl = upload_points(lib, L_enc)
    _, jr = upload_points(lib, R_enc)
    out = lib.fold(jl, jr, lib.upload(rt.scalars_to_bytes(xs)), rows, m)
    want = [cref.compress(cref.add(L_blob[i], cref.mul(R_blob[i], xs[i // m]))) for i in range(rows * m)]
    assert split48(lib.compress_jac(out, rows * m)) == want
    # group = m scalar-mul: one scalar per row
    out2 = lib.mul(jr, lib.upload(rt.scalars_to_bytes(xs)), rows * m, group=m)
    assert split48(lib.compress_jac(out2, rows * m)) == [cref.compress(cref.mul(R_blob[i], xs[i // m])) for i in range(rows * m)]


def case_msm(lib, cref, B, n, window=0, shared=False, seed=4, edge=False):
    rng = random.Random(seed * 1000 + n)
    nb = n if shared else B * n
    blobs, enc = rand_points(cref, rng, nb, with_identity=edge)
    if edge and nb >= 6:   # repeated bases and a base next to its own negation
        blobs[1], enc[1] = blobs[0], enc[0]
        blobs[3] = cref.neg(blobs[2]); enc[3] = cref.compress(blobs[3])
    ks = []
    for b in range(B):
        row = edge_scalars(rng, n) if edge and b % 2 == 0 else [rng.randrange(R) for _ in range(n)]
        if edge and b == 1:
            row = [0] * n
        if edge and b == 2 and n >= 4:
            row[2] = row[3] = 5      # 5*P + 5*(-P): the bucket meets P and -P
            row[0] = row[1] = 7      # same base twice with the same digit: doubling inside a bucket
        ks.append(row)
    aff, _ = upload_points(lib, enc)
    dk = lib.upload(rt.scalars_to_bytes([k for row in ks for k in row]))
    out = lib.msm_batched(aff, 0 if shared else n, dk, B, n, window)
    got = split48(lib.compress_jac(out, B))
    want = []
    for b in range(B):
        pts = blobs if shared else blobs[b * n:(b + 1) * n]
        want.append(cref.compress(cref.msm(pts, ks[b])))
    assert got == want, (B, n, window, shared)


def case_fixed(lib, cref, B, nb, window, seed=5, with_identity=False):
    """with_identity: the base vector also holds the identity - its whole table row is identities, which the shared
    inversions of the table build (Montgomery's trick per base / per 64-entry segment) must skip"""
    rng = random.Random(seed)
    blobs, enc = rand_points(cref, rng, nb, with_identity=with_identity)
    aff, _ = upload_points(lib, enc)
    table = lib.fixed_table(aff, nb, window)
    assert table.nbytes > 0
    ks = [edge_scalars(rng, nb) if b == 0 else [rng.randrange(R) for _ in range(nb)] for b in range(B)]
    dk = lib.upload(rt.scalars_to_bytes([k for row in ks for k in row]))
    out = lib.msm_fixed_batched(table, dk, B)
    want = [cref.msm(blobs, ks[b]) for b in range(B)]
    assert split48(lib.compress_jac(out, B)) == [cref.compress(w) for w in want]
    # accumulate=1 adds into the output: out + out
    lib.msm_fixed_batched(table, dk, B, accumulate=True, out=out)
    assert split48(lib.compress_jac(out, B)) == [cref.compress(cref.add(w, w)) for w in want]
    table.free()


def case_fr(lib, k, seed=6):
    rng = random.Random(seed)
    a = edge_scalars(rng, k)
    b = list(reversed(edge_scalars(rng, k)))
    da, db = lib.upload(rt.scalars_to_bytes(a)), lib.upload(rt.scalars_to_bytes(b))

    def ints(buf):
        raw = lib.download(buf, 32 * k)
        return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(k)]

    assert ints(lib.fr_op("add", da, db, k)) == [(x + y) % R for x, y in zip(a, b)]
    assert ints(lib.fr_op("sub", da, db, k)) == [(x - y) % R for x, y in zip(a, b)]
    assert ints(lib.fr_op("mul", da, db, k)) == [x * y % R for x, y in zip(a, b)]
    assert ints(lib.fr_op("inverse", da, None, k)) == [pow(x, -1, R) if x else 0 for x in a]


def case_msm_large(lib, cref, n, window, seed=8, slices=3):
    """One large MSM through the per-term (atomic) sort and the chunked window reduction, then the same
    value recomputed from window slices (the multi-GPU split) - all equal to the oracle."""
    rng = random.Random(seed * 77 + n)
    nuniq = min(n, 64)                                   # few distinct points, many terms: keeps the oracle fast
    ublobs, uenc = rand_points(cref, rng, nuniq, with_identity=True)
    idx = [rng.randrange(nuniq) for _ in range(n)]
    enc = [uenc[i] for i in idx]
    ks = [rng.randrange(R) for _ in range(n)]
    ks[0] = 0; ks[1] = R - 1
    aff, _ = upload_points(lib, enc)
    dk = lib.upload(rt.scalars_to_bytes(ks))
    # oracle: group the scalars per distinct base first (same value, far fewer scalar-muls)
    agg = [0] * nuniq
    for i, k in zip(idx, ks):
        agg[i] = (agg[i] + k) % R
    want = cref.compress(cref.msm(ublobs, agg))
    out = lib.msm_batched(aff, 0, dk, 1, n, window)
    assert split48(lib.compress_jac(out, 1)) == [want]
    c = window or int(lib.c.cpg_msm_pick_window(n))
    W = int(lib.c.cpg_msm_window_count(n, c))
    wsums = lib.alloc(W * rt.JAC)
    bounds = [W * i // slices for i in range(slices + 1)]
    for lo, hi in zip(bounds, bounds[1:]):
        if hi > lo:
            part = lib.alloc((hi - lo) * rt.JAC)
            lib.check(lib.c.cpg_g1_msm_window_sums(aff.ptr, dk.ptr, n, c, lo, hi, part.ptr), "cpg_g1_msm_window_sums")
            lib.check(lib.c.cpg_d2d(wsums.ptr + lo * rt.JAC, part.ptr, (hi - lo) * rt.JAC))
    res = lib.alloc(rt.JAC)
    lib.check(lib.c.cpg_g1_msm_combine_windows(wsums.ptr, c, res.ptr), "cpg_g1_msm_combine_windows")
    assert split48(lib.compress_jac(res, 1)) == [want]


def case_msm_skewed(lib, cref, n, window, seed=9):
    """Every term carries the same scalar: each window has ONE non-empty bucket holding all n terms - the run of a
    batched-affine thread overflows its scratch slots and must fall back to the XYZZ chain; a second instance uses two
    scalars (half / half), and identical bases throughout a bucket (doubling at every level of the pairwise tree)."""
    rng = random.Random(seed * 31 + n)
    blobs, enc = rand_points(cref, rng, 8, with_identity=False)
    k1, k2 = rng.randrange(R), rng.randrange(R)
    for bases_idx, ks in (([i % 8 for i in range(n)], [k1] * n),
                          ([i % 8 for i in range(n)], [k1 if i < n // 2 else k2 for i in range(n)]),
                          ([0] * n, [k1] * n)):
        aff, _ = upload_points(lib, [enc[i] for i in bases_idx])
        out = lib.msm_batched(aff, 0, lib.upload(rt.scalars_to_bytes(ks)), 1, n, window)
        agg = [0] * 8
        for i, k in zip(bases_idx, ks):
            agg[i] = (agg[i] + k) % R
        assert split48(lib.compress_jac(out, 1)) == [cref.compress(cref.msm(blobs, agg))], (n, window)

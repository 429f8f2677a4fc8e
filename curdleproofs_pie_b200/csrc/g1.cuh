// BLS12-381 G1 (y^2 = x^3 + 4 over Fq) point arithmetic for sm_100a, Montgomery-form coordinates.
//
// Replaces what the reference obtains from py_arkworks_bls12381's G1Point
// (/root/reference/curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-30): add/sub/neg/mul/eq,
// compressed (de)serialisation.  Three coordinate systems, each with one job:
//   Aff  (x, y)            96 B   bases of MSMs, fixed-base tables, decompress output; (0,0) = identity
//   Jac  (X, Y, Z)        144 B   general values and scalar-multiplication accumulators; Z = 0 = identity
//   Xyzz (X, Y, ZZ, ZZZ)  192 B   bucket accumulators (cheapest mixed add: 8M + 2S); ZZ = 0 = identity
// Everything here is CPG_HD: the same code runs in the kernels and, with the carry flag emulated
// (bigint.cuh), in the CPU-only unit tests (tests/host_seam).
#pragma once
#include "field.cuh"

namespace cpg {

struct Aff { Fq x, y; };
struct Jac { Fq X, Y, Z; };
struct Xyzz { Fq X, Y, ZZ, ZZZ; };

CPG_HD Fq fq_zero() { return Fq::zero(); }
CPG_HD Fq fq_one() { return Fq::one(); }
// curve constant b = 4 in Montgomery form
CPG_HD Fq fq_b4() { Fq o = fq_one(); Fq t = dbl(o); return dbl(t); }

CPG_HD bool is_inf(const Aff& p) { return p.x.is_zero() && p.y.is_zero(); }
CPG_HD bool is_inf(const Jac& p) { return p.Z.is_zero(); }
CPG_HD bool is_inf(const Xyzz& p) { return p.ZZ.is_zero(); }

CPG_HD Aff aff_inf() { Aff r; r.x = fq_zero(); r.y = fq_zero(); return r; }
CPG_HD Jac jac_inf() { Jac r; r.X = fq_one(); r.Y = fq_one(); r.Z = fq_zero(); return r; }
CPG_HD Xyzz xyzz_inf() { Xyzz r; r.X = fq_zero(); r.Y = fq_zero(); r.ZZ = fq_zero(); r.ZZZ = fq_zero(); return r; }

CPG_HD Jac to_jac(const Aff& p) {
    if (is_inf(p)) return jac_inf();
    Jac r; r.X = p.x; r.Y = p.y; r.Z = fq_one(); return r;
}
CPG_HD Xyzz to_xyzz(const Aff& p) {
    if (is_inf(p)) return xyzz_inf();
    Xyzz r; r.X = p.x; r.Y = p.y; r.ZZ = fq_one(); r.ZZZ = fq_one(); return r;
}
CPG_HD Aff neg(const Aff& p) { Aff r; r.x = p.x; r.y = neg(p.y); return r; }
CPG_HD Jac neg(const Jac& p) { Jac r; r.X = p.X; r.Y = neg(p.Y); r.Z = p.Z; return r; }
// conditional negation (used for signed digits)
CPG_HD Aff cneg(const Aff& p, bool s) { Aff r; r.x = p.x; r.y = s ? neg(p.y) : p.y; return r; }

// ---------------------------------------------------------------- Jacobian ---
// dbl-2009-l (a = 0): 2M + 5S
CPG_HD Jac jac_dbl(const Jac& p) {
    if (is_inf(p)) return p;
    Fq A = sqr(p.X), B = sqr(p.Y), C = sqr(B);
    Fq t = add(p.X, B);
    Fq D = sub(sub(sqr(t), A), C); D = dbl(D);
    Fq E = add(dbl(A), A);
    Fq F = sqr(E);
    Jac r;
    r.Z = dbl(mul(p.Y, p.Z));
    r.X = sub(F, dbl(D));
    Fq C8 = dbl(dbl(dbl(C)));
    r.Y = sub(mul(E, sub(D, r.X)), C8);
    return r;
}
// add-2007-bl: 11M + 5S, complete via explicit special cases
CPG_HD Jac jac_add(const Jac& p, const Jac& q) {
    if (is_inf(p)) return q;
    if (is_inf(q)) return p;
    Fq Z1Z1 = sqr(p.Z), Z2Z2 = sqr(q.Z);
    Fq U1 = mul(p.X, Z2Z2), U2 = mul(q.X, Z1Z1);
    Fq S1 = mul(mul(p.Y, q.Z), Z2Z2), S2 = mul(mul(q.Y, p.Z), Z1Z1);
    Fq H = sub(U2, U1), rr = sub(S2, S1);
    if (H.is_zero()) {
        if (rr.is_zero()) return jac_dbl(p);
        return jac_inf();
    }
    rr = dbl(rr);
    Fq I = sqr(dbl(H));
    Fq J = mul(H, I);
    Fq V = mul(U1, I);
    Jac r;
    r.X = sub(sub(sqr(rr), J), dbl(V));
    r.Y = sub(mul(rr, sub(V, r.X)), dbl(mul(S1, J)));
    r.Z = mul(sub(sub(sqr(add(p.Z, q.Z)), Z1Z1), Z2Z2), H);
    return r;
}
// madd-2007-bl: 7M + 4S
CPG_HD Jac jac_add_mixed(const Jac& p, const Aff& q) {
    if (is_inf(q)) return p;
    if (is_inf(p)) return to_jac(q);
    Fq Z1Z1 = sqr(p.Z);
    Fq U2 = mul(q.x, Z1Z1);
    Fq S2 = mul(mul(q.y, p.Z), Z1Z1);
    Fq H = sub(U2, p.X), rr = sub(S2, p.Y);
    if (H.is_zero()) {
        if (rr.is_zero()) return jac_dbl(p);
        return jac_inf();
    }
    rr = dbl(rr);
    Fq HH = sqr(H);
    Fq I = dbl(dbl(HH));
    Fq J = mul(H, I);
    Fq V = mul(p.X, I);
    Jac r;
    r.X = sub(sub(sqr(rr), J), dbl(V));
    r.Y = sub(mul(rr, sub(V, r.X)), dbl(mul(p.Y, J)));
    r.Z = sub(sub(sqr(add(p.Z, H)), Z1Z1), HH);
    return r;
}
CPG_HD bool jac_eq(const Jac& p, const Jac& q) {
    bool ip = is_inf(p), iq = is_inf(q);
    if (ip || iq) return ip && iq;
    Fq Z1Z1 = sqr(p.Z), Z2Z2 = sqr(q.Z);
    if (mul(p.X, Z2Z2) != mul(q.X, Z1Z1)) return false;
    return mul(mul(p.Y, q.Z), Z2Z2) == mul(mul(q.Y, p.Z), Z1Z1);
}
// One Fq inversion.  Batched callers use batch_to_affine-style kernels instead.
CPG_HD Aff jac_to_aff(const Jac& p) {
    if (is_inf(p)) return aff_inf();
    Fq zi = fq_inv(p.Z);
    Fq zi2 = sqr(zi);
    Aff r; r.x = mul(p.X, zi2); r.y = mul(p.Y, mul(zi2, zi));
    return r;
}

// -------------------------------------------------------------------- XYZZ ---
// mdbl-2008-s-1 (from affine): 4M... used when a bucket meets the same point twice
CPG_HD Xyzz xyzz_dbl_aff(const Aff& p) {
    if (is_inf(p) || p.y.is_zero()) return xyzz_inf();
    Fq U = dbl(p.y), V = sqr(U), W = mul(U, V), S = mul(p.x, V);
    Fq X2 = sqr(p.x);
    Fq M = add(dbl(X2), X2);
    Xyzz r;
    r.X = sub(sqr(M), dbl(S));
    r.Y = sub(mul(M, sub(S, r.X)), mul(W, p.y));
    r.ZZ = V; r.ZZZ = W;
    return r;
}
// dbl-2008-s-1: 6M + 3S
CPG_HD Xyzz xyzz_dbl(const Xyzz& p) {
    if (is_inf(p)) return p;
    Fq U = dbl(p.Y), V = sqr(U), W = mul(U, V), S = mul(p.X, V);
    Fq X2 = sqr(p.X);
    Fq M = add(dbl(X2), X2);
    Xyzz r;
    r.X = sub(sqr(M), dbl(S));
    r.Y = sub(mul(M, sub(S, r.X)), mul(W, p.Y));
    r.ZZ = mul(V, p.ZZ); r.ZZZ = mul(W, p.ZZZ);
    return r;
}
// madd-2008-s: 8M + 2S.  The bucket-accumulation workhorse.
CPG_HD Xyzz xyzz_add_mixed(const Xyzz& p, const Aff& q) {
    if (is_inf(q)) return p;
    if (is_inf(p)) return to_xyzz(q);
    Fq U2 = mul(q.x, p.ZZ), S2 = mul(q.y, p.ZZZ);
    Fq P = sub(U2, p.X), R = sub(S2, p.Y);
    if (P.is_zero()) {
        if (R.is_zero()) return xyzz_dbl_aff(q);
        return xyzz_inf();
    }
    Fq PP = sqr(P), PPP = mul(P, PP), Q = mul(p.X, PP);
    Xyzz r;
    r.X = sub(sub(sqr(R), PPP), dbl(Q));
    r.Y = sub(mul(R, sub(Q, r.X)), mul(p.Y, PPP));
    r.ZZ = mul(p.ZZ, PP); r.ZZZ = mul(p.ZZZ, PPP);
    return r;
}
// add-2008-s: 12M + 2S
CPG_HD Xyzz xyzz_add(const Xyzz& p, const Xyzz& q) {
    if (is_inf(q)) return p;
    if (is_inf(p)) return q;
    Fq U1 = mul(p.X, q.ZZ), U2 = mul(q.X, p.ZZ);
    Fq S1 = mul(p.Y, q.ZZZ), S2 = mul(q.Y, p.ZZZ);
    Fq P = sub(U2, U1), R = sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) return xyzz_dbl(p);
        return xyzz_inf();
    }
    Fq PP = sqr(P), PPP = mul(P, PP), Q = mul(U1, PP);
    Xyzz r;
    r.X = sub(sub(sqr(R), PPP), dbl(Q));
    r.Y = sub(mul(R, sub(Q, r.X)), mul(S1, PPP));
    r.ZZ = mul(mul(p.ZZ, q.ZZ), PP); r.ZZZ = mul(mul(p.ZZZ, q.ZZZ), PPP);
    return r;
}
// (X, Y, ZZ, ZZZ) -> Jacobian with Z = ZZ*ZZZ... no inversion:  x = X/ZZ, y = Y/ZZZ and
// ZZ^3 = ZZZ^2, so with Z' = ZZZ/ZZ... we avoid the division by picking Z' = ZZ*ZZZ (= Z^5):
//   X' = x Z'^2 = X ZZ ZZZ^2,   Y' = y Z'^3 = Y ZZ^3 ZZZ^2 = Y ZZZ^4.   5M + 2S
CPG_HD Jac xyzz_to_jac(const Xyzz& p) {
    if (is_inf(p)) return jac_inf();
    Fq z3sq = sqr(p.ZZZ);
    Jac r;
    r.X = mul(mul(p.X, p.ZZ), z3sq);
    r.Y = mul(p.Y, sqr(z3sq));
    r.Z = mul(p.ZZ, p.ZZZ);
    return r;
}
CPG_HD Aff xyzz_to_aff(const Xyzz& p) {
    if (is_inf(p)) return aff_inf();
    Fq i = fq_inv(mul(p.ZZ, p.ZZZ));
    Aff r; r.x = mul(p.X, mul(i, p.ZZZ)); r.y = mul(p.Y, mul(i, p.ZZ));
    return r;
}

// ---------------------------------------------------- scalar multiplication ---
// k * P for a canonical 255-bit scalar (8 x u32 little-endian, NOT Montgomery), signed 4-bit
// windows over a table {1..8}P.  Replaces G1Point.__mul__ (stub :10); used by the element-wise
// kernels (vector scalar-mul / fold), never inside the MSMs.
CPG_HD Jac jac_mul(const Jac& p, const uint32_t* k) {
    if (is_inf(p)) return jac_inf();
    Jac tbl[8];
    tbl[0] = p;
    tbl[1] = jac_dbl(p);
    for (int i = 2; i < 8; i++) tbl[i] = jac_add(tbl[i - 1], p);
    // signed recoding, local form: k' = k + sum_w 8*16^w (w < 64); digit_w = nibble_w(k') - 8.
    // k < r < 0.91 * 2^255 so k' < 2^256 (see DESIGN.md "signed digits").
    uint32_t kp[8];
    kp[0] = add_cc(k[0], 0x88888888u);
    for (int i = 1; i < 7; i++) kp[i] = addc_cc(k[i], 0x88888888u);
    kp[7] = addc(k[7], 0x88888888u);
    Jac acc = jac_inf();
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { acc = jac_dbl(acc); acc = jac_dbl(acc); acc = jac_dbl(acc); acc = jac_dbl(acc); }
        int d = (int)((kp[w >> 3] >> ((w & 7) * 4)) & 15u) - 8;
        if (d > 0) acc = jac_add(acc, tbl[d - 1]);
        else if (d < 0) acc = jac_add(acc, neg(tbl[-d - 1]));
    }
    return acc;
}

// GLV: on G1 the endomorphism phi(x, y) = (beta x, y) acts as multiplication by lambda = z^2 - 1
// (lambda^2 + lambda + 1 = 0 mod r, beta^3 = 1 in Fq; pair checked against the oracle: phi(G) = lambda G).
// With k = k1 + k2 lambda as INTEGERS (k1 = k mod lambda, k2 = k div lambda <= lambda + 1, both < 2^128; the host
// splits, prove.inl::glv_split)  k P = k1 P + k2 phi(P): 128 doublings + 2 x 33 signed 4-bit digits instead of
// 255 doublings + 64 digits (~2040 vs ~2860 Fq products).  Valid for P in the r-order subgroup only - as
// arkworks' own G1 scalar multiplication (GLV as well); behaviour on points outside G1 is unpinned anyway.
#define CPG_GLV_BETA_INIT {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu, \
                           0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u}
static const uint32_t H_GLV_BETA[12] = CPG_GLV_BETA_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t D_GLV_BETA[12] = CPG_GLV_BETA_INIT;
#endif
CPG_HD Jac jac_mul_glv(const Jac& p, const uint32_t* k12 /* k1[4] | k2[4], little-endian words */) {
    if (is_inf(p)) return jac_inf();
    Jac tbl[8];
    tbl[0] = p;
    tbl[1] = jac_dbl(p);
    for (int i = 2; i < 8; i++) tbl[i] = jac_add(tbl[i - 1], p);
    Fq beta;
    for (int i = 0; i < 12; i++) beta.l[i] = CPG_SEL(GLV_BETA)[i];
    // signed recoding, local form, of both halves: k' = k + sum_{w<32} 8*16^w < 2^129; digit_w = nibble_w(k') - 8,
    // top digit (w = 32) = k' >> 128 in {0, 1}
    uint32_t a[5], b[5];
    a[0] = add_cc(k12[0], 0x88888888u); a[1] = addc_cc(k12[1], 0x88888888u); a[2] = addc_cc(k12[2], 0x88888888u);
    a[3] = addc_cc(k12[3], 0x88888888u); a[4] = addc(0, 0);
    b[0] = add_cc(k12[4], 0x88888888u); b[1] = addc_cc(k12[5], 0x88888888u); b[2] = addc_cc(k12[6], 0x88888888u);
    b[3] = addc_cc(k12[7], 0x88888888u); b[4] = addc(0, 0);
    Jac acc = jac_inf();
    for (int w = 32; w >= 0; w--) {
        if (w != 32) { acc = jac_dbl(acc); acc = jac_dbl(acc); acc = jac_dbl(acc); acc = jac_dbl(acc); }
        int d1 = w == 32 ? (int)a[4] : (int)((a[w >> 3] >> ((w & 7) * 4)) & 15u) - 8;
        int d2 = w == 32 ? (int)b[4] : (int)((b[w >> 3] >> ((w & 7) * 4)) & 15u) - 8;
        if (d1 > 0) acc = jac_add(acc, tbl[d1 - 1]);
        else if (d1 < 0) acc = jac_add(acc, neg(tbl[-d1 - 1]));
        if (d2) {
            Jac q = tbl[(d2 < 0 ? -d2 : d2) - 1];
            q.X = mul(q.X, beta);                       // phi in Jacobian coordinates: x = X / Z^2
            acc = jac_add(acc, d2 < 0 ? neg(q) : q);
        }
    }
    return acc;
}

// ------------------------------------------------------------ serialisation ---
// 48-byte ZCash/IETF compressed form; replaces to_compressed_bytes (stub :30).
CPG_HD void aff_compress(const Aff& p, uint8_t* out) {
    if (is_inf(p)) {
        out[0] = 0xc0;
        for (int i = 1; i < 48; i++) out[i] = 0;
        return;
    }
    Fq x = from_mont(p.x), y = from_mont(p.y);
    for (int i = 0; i < 12; i++) {
        uint32_t w = x.l[11 - i];
        out[4 * i] = (uint8_t)(w >> 24); out[4 * i + 1] = (uint8_t)(w >> 16);
        out[4 * i + 2] = (uint8_t)(w >> 8); out[4 * i + 3] = (uint8_t)w;
    }
    out[0] |= 0x80;
    if (fq_is_lex_largest(y)) out[0] |= 0x20;
}
// Returns 0 on success, else an error code (1 no compression flag, 2 bad infinity encoding,
// 3 x >= p, 4 not on curve, 5 not in the r-order subgroup).  Replaces from_compressed_bytes
// (checked, stub :19) and from_compressed_bytes_unchecked (stub :22).
CPG_HD int aff_decompress(const uint8_t* in, bool check_subgroup, Aff* out) {
    uint8_t flags = in[0];
    *out = aff_inf();
    if (!(flags & 0x80)) return 1;
    Fq x;
    for (int i = 0; i < 12; i++) {
        const uint8_t* b = in + 4 * (11 - i);
        x.l[i] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
    }
    x.l[11] &= 0x1fffffffu;
    if (flags & 0x40) {
        if (!x.is_zero() || (flags & 0x20)) return 2;
        return 0;
    }
    if (!is_canonical(x)) return 3;
    Fq xm = to_mont(x);
    Fq rhs = add(mul(sqr(xm), xm), fq_b4());
    Fq y = fq_sqrt_candidate(rhs);
    if (sqr(y) != rhs) return 4;
    bool big = fq_is_lex_largest(from_mont(y));
    if (big != ((flags & 0x20) != 0)) y = neg(y);
    out->x = xm; out->y = y;
    if (check_subgroup) {
        // r * P == O  (r itself as the multiplier: jac_mul's recoding needs k < 2^255 * 0.93, r fits)
        const uint32_t rr[8] = CPG_FR_P_INIT;
        Jac t = jac_mul(to_jac(*out), rr);
        if (!is_inf(t)) { *out = aff_inf(); return 5; }
    }
    return 0;
}

}  // namespace cpg

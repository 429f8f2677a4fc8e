/* ORACLE (test infrastructure, never on the product path).
 *
 * Plain-C CPU restatement of the BLS12-381 G1 arithmetic the reference obtains from the
 * external wheel py_arkworks_bls12381 0.3.5 (/root/reference/curdleproofs/pyproject.toml:10;
 * stub curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-30), plus the reference's
 * own Keccak-f[1600]/STROBE-128 permutation (merlin_transcripts/merlin_transcripts/keccak.py:56-66).
 * The algorithms are the ones SURVEY.md 3.4 attributes to arkworks:
 *   - G1Point * Scalar : MSB-first double-and-add in Jacobian coordinates
 *   - multiexp_unchecked: Pippenger bucket method with arkworks' window rule
 *   - compute_MSM (cp/msm_accumulator.py:6-12): the naive loop  acc += base*scalar
 * Checked against oracle/bls12381_py.py (Python big integers) and the reference's KATs in
 * tests/test_oracle_*.py.  Only tests/, smoke() and bench.py's CPU legs may link this.
 *
 * Wire format of a point ("blob"): 144 bytes = X | Y | Z, each 48-byte little-endian
 * canonical integers < p (Jacobian); Z == 0 encodes the identity.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define NL 6
typedef struct { u64 l[NL]; } fp;

static const fp FP_P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                         0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const u64 FP_INV = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */
/* R = 2^384 mod p, R2 = 2^768 mod p */
static const fp FP_R = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                         0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
static const fp FP_R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                          0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};

static int fp_geq(const fp *a, const fp *b) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static int fp_is_zero(const fp *a) {
    u64 t = 0;
    for (int i = 0; i < NL; i++) t |= a->l[i];
    return t == 0;
}
static int fp_eq(const fp *a, const fp *b) {
    u64 t = 0;
    for (int i = 0; i < NL; i++) t |= a->l[i] ^ b->l[i];
    return t == 0;
}
static u64 raw_add(fp *r, const fp *a, const fp *b) {
    u128 c = 0;
    for (int i = 0; i < NL; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static u64 raw_sub(fp *r, const fp *a, const fp *b) {
    u64 borrow = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        r->l[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static void fp_add(fp *r, const fp *a, const fp *b) {
    fp t; raw_add(&t, a, b);           /* p < 2^381: no carry out */
    if (fp_geq(&t, &FP_P)) raw_sub(&t, &t, &FP_P);
    *r = t;
}
static void fp_sub(fp *r, const fp *a, const fp *b) {
    fp t; if (raw_sub(&t, a, b)) raw_add(&t, &t, &FP_P);
    *r = t;
}
static void fp_neg(fp *r, const fp *a) {
    if (fp_is_zero(a)) { *r = *a; return; }
    raw_sub(r, &FP_P, a);
}
/* Montgomery product a*b/2^384 mod p (coarsely integrated operand scanning) */
static inline __attribute__((always_inline)) void fp_mul(fp *r, const fp *a, const fp *b) {
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
#pragma GCC unroll 6
    for (int i = 0; i < NL; i++) {
        const u64 bi = b->l[i];
        u128 c;
        c = (u128)a->l[0] * bi + t0; t0 = (u64)c; c >>= 64;
        c += (u128)a->l[1] * bi + t1; t1 = (u64)c; c >>= 64;
        c += (u128)a->l[2] * bi + t2; t2 = (u64)c; c >>= 64;
        c += (u128)a->l[3] * bi + t3; t3 = (u64)c; c >>= 64;
        c += (u128)a->l[4] * bi + t4; t4 = (u64)c; c >>= 64;
        c += (u128)a->l[5] * bi + t5; t5 = (u64)c; c >>= 64;
        c += t6; t6 = (u64)c;                    /* total < 2^(384+64): no further carry */
        const u64 m = t0 * FP_INV;
        c = ((u128)m * FP_P.l[0] + t0) >> 64;
        c += (u128)m * FP_P.l[1] + t1; t0 = (u64)c; c >>= 64;
        c += (u128)m * FP_P.l[2] + t2; t1 = (u64)c; c >>= 64;
        c += (u128)m * FP_P.l[3] + t3; t2 = (u64)c; c >>= 64;
        c += (u128)m * FP_P.l[4] + t4; t3 = (u64)c; c >>= 64;
        c += (u128)m * FP_P.l[5] + t5; t4 = (u64)c; c >>= 64;
        c += t6; t5 = (u64)c; t6 = (u64)(c >> 64);
    }
    fp out = {{t0, t1, t2, t3, t4, t5}};
    if (t6 || fp_geq(&out, &FP_P)) raw_sub(&out, &out, &FP_P);
    *r = out;
}
static void fp_sqr(fp *r, const fp *a) { fp_mul(r, a, a); }
static void fp_to_mont(fp *r, const fp *a) { fp_mul(r, a, &FP_R2); }
static void fp_from_mont(fp *r, const fp *a) { fp one = {{1, 0, 0, 0, 0, 0}}; fp_mul(r, a, &one); }
/* r = a^e, e given as little-endian 64-bit words */
static void fp_pow(fp *r, const fp *a, const u64 *e, int nwords) {
    fp acc = FP_R;
    int started = 0;
    for (int i = nwords * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) { fp_mul(&acc, &acc, a); started = 1; }
    }
    *r = acc;
}
static void fp_inv(fp *r, const fp *a) {
    fp e = FP_P; e.l[0] -= 2; fp_pow(r, a, e.l, NL);
}
static int fp_sqrt(fp *r, const fp *a) { /* p = 3 mod 4 */
    fp e; fp one = {{1, 0, 0, 0, 0, 0}};
    raw_add(&e, &FP_P, &one);
    for (int i = 0; i < NL; i++) e.l[i] = (e.l[i] >> 2) | (i + 1 < NL ? e.l[i + 1] << 62 : 0);
    fp y; fp_pow(&y, a, e.l, NL);
    fp y2; fp_sqr(&y2, &y);
    *r = y;
    return fp_eq(&y2, a);
}
static void fp_from_le(fp *r, const uint8_t *b) {
    for (int i = 0; i < NL; i++) { u64 w = 0; for (int k = 7; k >= 0; k--) w = (w << 8) | b[8 * i + k]; r->l[i] = w; }
}
static void fp_to_le(uint8_t *b, const fp *a) {
    for (int i = 0; i < NL; i++) for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(a->l[i] >> (8 * k));
}

/* ---- G1, Jacobian, Montgomery-form coordinates internally ---- */
typedef struct { fp X, Y, Z; } g1;

static void g1_set_inf(g1 *r) { memset(r, 0, sizeof *r); r->X = FP_R; r->Y = FP_R; }
static int g1_is_inf(const g1 *a) { return fp_is_zero(&a->Z); }

static void g1_load(g1 *r, const uint8_t *blob) {
    fp t;
    fp_from_le(&t, blob); fp_to_mont(&r->X, &t);
    fp_from_le(&t, blob + 48); fp_to_mont(&r->Y, &t);
    fp_from_le(&t, blob + 96); fp_to_mont(&r->Z, &t);
}
static void g1_store(uint8_t *blob, const g1 *a) {
    fp t;
    fp_from_mont(&t, &a->X); fp_to_le(blob, &t);
    fp_from_mont(&t, &a->Y); fp_to_le(blob + 48, &t);
    fp_from_mont(&t, &a->Z); fp_to_le(blob + 96, &t);
}
static void g1_dbl(g1 *r, const g1 *a) {
    if (g1_is_inf(a) || fp_is_zero(&a->Y)) { g1_set_inf(r); return; }
    fp A, B, C, D, E, F, t, X3, Y3, Z3;
    fp_sqr(&A, &a->X); fp_sqr(&B, &a->Y); fp_sqr(&C, &B);
    fp_add(&t, &a->X, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_add(&D, &t, &t);
    fp_add(&E, &A, &A); fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp_sub(&X3, &F, &D); fp_sub(&X3, &X3, &D);
    fp_sub(&t, &D, &X3); fp_mul(&Y3, &E, &t);
    fp_add(&C, &C, &C); fp_add(&C, &C, &C); fp_add(&C, &C, &C);
    fp_sub(&Y3, &Y3, &C);
    fp_mul(&Z3, &a->Y, &a->Z); fp_add(&Z3, &Z3, &Z3);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1_add(g1 *r, const g1 *a, const g1 *b) {
    if (g1_is_inf(a)) { *r = *b; return; }
    if (g1_is_inf(b)) { *r = *a; return; }
    fp Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
    fp_sqr(&Z1Z1, &a->Z); fp_sqr(&Z2Z2, &b->Z);
    fp_mul(&U1, &a->X, &Z2Z2); fp_mul(&U2, &b->X, &Z1Z1);
    fp_mul(&S1, &a->Y, &b->Z); fp_mul(&S1, &S1, &Z2Z2);
    fp_mul(&S2, &b->Y, &a->Z); fp_mul(&S2, &S2, &Z1Z1);
    if (fp_eq(&U1, &U2)) {
        if (fp_eq(&S1, &S2)) { g1_dbl(r, a); return; }
        g1_set_inf(r); return;
    }
    fp_sub(&H, &U2, &U1);
    fp_add(&I, &H, &H); fp_sqr(&I, &I);
    fp_mul(&J, &H, &I);
    fp_sub(&rr, &S2, &S1); fp_add(&rr, &rr, &rr);
    fp_mul(&V, &U1, &I);
    fp_sqr(&X3, &rr); fp_sub(&X3, &X3, &J); fp_sub(&X3, &X3, &V); fp_sub(&X3, &X3, &V);
    fp_sub(&t, &V, &X3); fp_mul(&Y3, &rr, &t);
    fp_mul(&t, &S1, &J); fp_add(&t, &t, &t); fp_sub(&Y3, &Y3, &t);
    fp_add(&Z3, &a->Z, &b->Z); fp_sqr(&Z3, &Z3); fp_sub(&Z3, &Z3, &Z1Z1); fp_sub(&Z3, &Z3, &Z2Z2);
    fp_mul(&Z3, &Z3, &H);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1_neg(g1 *r, const g1 *a) { *r = *a; fp_neg(&r->Y, &a->Y); }
static int g1_eq(const g1 *a, const g1 *b) {
    int ia = g1_is_inf(a), ib = g1_is_inf(b);
    if (ia || ib) return ia && ib;
    fp Z1Z1, Z2Z2, l, rr;
    fp_sqr(&Z1Z1, &a->Z); fp_sqr(&Z2Z2, &b->Z);
    fp_mul(&l, &a->X, &Z2Z2); fp_mul(&rr, &b->X, &Z1Z1);
    if (!fp_eq(&l, &rr)) return 0;
    fp_mul(&l, &a->Y, &b->Z); fp_mul(&l, &l, &Z2Z2);
    fp_mul(&rr, &b->Y, &a->Z); fp_mul(&rr, &rr, &Z1Z1);
    return fp_eq(&l, &rr);
}
/* k: 32-byte little-endian, any value < 2^256 (callers pass canonical < r) */
static void g1_mul(g1 *r, const g1 *a, const uint8_t *k) {
    g1 acc; g1_set_inf(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) g1_dbl(&acc, &acc);
        if ((k[i / 8] >> (i % 8)) & 1) { g1_add(&acc, &acc, a); started = 1; }
    }
    *r = acc;
}
static void g1_to_affine(fp *x, fp *y, const g1 *a) {
    fp zi, zi2;
    fp_inv(&zi, &a->Z); fp_sqr(&zi2, &zi);
    fp_mul(x, &a->X, &zi2); fp_mul(&zi2, &zi2, &zi); fp_mul(y, &a->Y, &zi2);
}

static const uint8_t FR_R_LE[32] = {0x01, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0x02, 0xa4, 0xbd, 0x53,
                                    0x05, 0xd8, 0xa1, 0x09, 0x08, 0xd8, 0x39, 0x33, 0x48, 0x7d, 0x9d, 0x29, 0x53, 0xa7, 0xed, 0x73};

/* ---- exported flat C ABI ---- */
void ref_g1_generator(uint8_t *out) {
    static const uint8_t gx[48] = {0xbb, 0xc6, 0x22, 0xdb, 0x0a, 0xf0, 0x3a, 0xfb, 0xef, 0x1a, 0x7a, 0xf9, 0x3f, 0xe8, 0x55, 0x6c,
                                   0x58, 0xac, 0x1b, 0x17, 0x3f, 0x3a, 0x4e, 0xa1, 0x05, 0xb9, 0x74, 0x97, 0x4f, 0x8c, 0x68, 0xc3,
                                   0x0f, 0xac, 0xa9, 0x4f, 0x8c, 0x63, 0x95, 0x26, 0x94, 0xd7, 0x97, 0x31, 0xa7, 0xd3, 0xf1, 0x17};
    static const uint8_t gy[48] = {0xe1, 0xe7, 0xc5, 0x46, 0x29, 0x23, 0xaa, 0x0c, 0xe4, 0x8a, 0x88, 0xa2, 0x44, 0xc7, 0x3c, 0xd0,
                                   0xed, 0xb3, 0x04, 0x2c, 0xcb, 0x18, 0xdb, 0x00, 0xf6, 0x0a, 0xd0, 0xd5, 0x95, 0xe0, 0xf5, 0xfc,
                                   0xe4, 0x8a, 0x1d, 0x74, 0xed, 0x30, 0x9e, 0xa0, 0xf1, 0xa0, 0xaa, 0xe3, 0x81, 0xf4, 0xb3, 0x08};
    memset(out, 0, 144); memcpy(out, gx, 48); memcpy(out + 48, gy, 48); out[96] = 1;
}
void ref_g1_identity(uint8_t *out) { memset(out, 0, 144); out[0] = 1; out[48] = 1; }
void ref_g1_add(const uint8_t *a, const uint8_t *b, uint8_t *out) { g1 x, y, z; g1_load(&x, a); g1_load(&y, b); g1_add(&z, &x, &y); g1_store(out, &z); }
void ref_g1_sub(const uint8_t *a, const uint8_t *b, uint8_t *out) { g1 x, y, z; g1_load(&x, a); g1_load(&y, b); g1_neg(&y, &y); g1_add(&z, &x, &y); g1_store(out, &z); }
void ref_g1_neg(const uint8_t *a, uint8_t *out) { g1 x; g1_load(&x, a); g1_neg(&x, &x); g1_store(out, &x); }
int ref_g1_eq(const uint8_t *a, const uint8_t *b) { g1 x, y; g1_load(&x, a); g1_load(&y, b); return g1_eq(&x, &y); }
void ref_g1_mul(const uint8_t *a, const uint8_t *k, uint8_t *out) { g1 x, z; g1_load(&x, a); g1_mul(&z, &x, k); g1_store(out, &z); }

void ref_g1_compress(const uint8_t *a, uint8_t *out) {
    g1 x; g1_load(&x, a);
    if (g1_is_inf(&x)) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fp ax, ay, t, half;
    g1_to_affine(&ax, &ay, &x);
    fp_from_mont(&t, &ax);
    uint8_t le[48]; fp_to_le(le, &t);
    for (int i = 0; i < 48; i++) out[i] = le[47 - i];
    out[0] |= 0x80;
    fp_from_mont(&t, &ay);
    /* y > (p-1)/2  <=>  2y > p-1  <=>  2y >= p (p odd)  <=>  y >= p - y with y != 0 */
    fp_neg(&half, &t); /* half = p - y (canonical ints here, fp_neg is plain subtraction) */
    if (!fp_is_zero(&t) && fp_geq(&t, &half) && !fp_eq(&t, &half)) out[0] |= 0x20;
}
int ref_g1_decompress(const uint8_t *in, int check_subgroup, uint8_t *out) {
    uint8_t flags = in[0];
    if (!(flags & 0x80)) return 1;
    uint8_t le[48];
    for (int i = 0; i < 48; i++) le[i] = in[47 - i];
    le[47] &= 0x1f;
    fp x; fp_from_le(&x, le);
    if (flags & 0x40) {
        if (!fp_is_zero(&x) || (flags & 0x20)) return 2;
        ref_g1_identity(out); return 0;
    }
    if (fp_geq(&x, &FP_P)) return 3;
    fp xm, rhs, four = {{4, 0, 0, 0, 0, 0}}, y, yc, ny;
    fp_to_mont(&xm, &x); fp_to_mont(&four, &four);
    fp_sqr(&rhs, &xm); fp_mul(&rhs, &rhs, &xm); fp_add(&rhs, &rhs, &four);
    if (!fp_sqrt(&y, &rhs)) return 4;
    fp_from_mont(&yc, &y); fp_neg(&ny, &yc);
    int big = !fp_is_zero(&yc) && fp_geq(&yc, &ny) && !fp_eq(&yc, &ny);
    if (big != !!(flags & 0x20)) fp_neg(&y, &y);
    g1 pt; pt.X = xm; pt.Y = y; pt.Z = FP_R;
    if (check_subgroup) { g1 t; g1_mul(&t, &pt, FR_R_LE); if (!g1_is_inf(&t)) return 5; }
    g1_store(out, &pt);
    return 0;
}
/* the reference's compute_MSM loop, cp/msm_accumulator.py:6-12 */
void ref_g1_msm_naive(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out) {
    g1 acc, p, t; g1_set_inf(&acc);
    for (size_t i = 0; i < n; i++) { g1_load(&p, pts + 144 * i); g1_mul(&t, &p, scalars + 32 * i); g1_add(&acc, &acc, &t); }
    g1_store(out, &acc);
}
static unsigned get_bits(const uint8_t *k, int lo, int c) {
    unsigned v = 0;
    for (int b = 0; b < c; b++) { int i = lo + b; if (i < 256 && ((k[i / 8] >> (i % 8)) & 1)) v |= 1u << b; }
    return v;
}
/* bucket method with the window rule SURVEY 3.4 attributes to arkworks */
void ref_g1_msm_pippenger(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out) {
    g1 total; g1_set_inf(&total);
    if (n == 0) { g1_store(out, &total); return; }
    int lg = 0; while (((size_t)1 << lg) < n) lg++;
    int c = n < 32 ? 3 : (lg * 69) / 100 + 2;
    int nwin = (255 + c - 1) / c;
    g1 *P = malloc(n * sizeof(g1));
    g1 *bk = malloc(((size_t)1 << c) * sizeof(g1));
    for (size_t i = 0; i < n; i++) g1_load(&P[i], pts + 144 * i);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) g1_dbl(&total, &total);
        for (unsigned d = 0; d < (1u << c); d++) g1_set_inf(&bk[d]);
        for (size_t i = 0; i < n; i++) { unsigned d = get_bits(scalars + 32 * i, w * c, c); if (d) g1_add(&bk[d], &bk[d], &P[i]); }
        g1 run, acc; g1_set_inf(&run); g1_set_inf(&acc);
        for (unsigned d = (1u << c) - 1; d >= 1; d--) { g1_add(&run, &run, &bk[d]); g1_add(&acc, &acc, &run); }
        g1_add(&total, &total, &acc);
    }
    free(P); free(bk);
    g1_store(out, &total);
}
/* batch helpers (amortise the ctypes call) */
void ref_g1_mul_batch(const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out) { for (size_t i = 0; i < n; i++) ref_g1_mul(pts + 144 * i, scalars + 32 * i, out + 144 * i); }
void ref_g1_compress_batch(const uint8_t *pts, size_t n, uint8_t *out) { for (size_t i = 0; i < n; i++) ref_g1_compress(pts + 144 * i, out + 48 * i); }
void ref_g1_decompress_batch(const uint8_t *in, size_t n, int check, uint8_t *out, uint8_t *ok) { for (size_t i = 0; i < n; i++) ok[i] = (uint8_t)(ref_g1_decompress(in + 48 * i, check, out + 144 * i) == 0); }

/* ---- Keccak-f[1600] (merlin_transcripts/merlin_transcripts/keccak.py:56-66 calls the same permutation) ---- */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static u64 rol(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
void ref_keccak_f1600(uint8_t *state) {
    u64 A[25];
    for (int i = 0; i < 25; i++) { u64 w = 0; for (int k = 7; k >= 0; k--) w = (w << 8) | state[8 * i + k]; A[i] = w; }
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol(A[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= KRC[rnd];
    }
    for (int i = 0; i < 25; i++) for (int k = 0; k < 8; k++) state[8 * i + k] = (uint8_t)(A[i] >> (8 * k));
}

// BLS12-381 base field Fq (12 x u32) and scalar field Fr (8 x u32), Montgomery form.
// Constants verified numerically in SURVEY.md A.1.
#pragma once
#include "bigint.cuh"

namespace cpg {

#if defined(__CUDACC__)
#define CPG_CONST_DECL __device__ __constant__
#else
#define CPG_CONST_DECL static const
#endif

// Device code reads the __constant__ copies (IMAD takes a constant-bank operand directly);
// host code (unit tests of the same algorithms) reads the plain arrays.
#define CPG_FQ_P_INIT {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, \
                       0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define CPG_FQ_R_INIT {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, \
                       0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}
#define CPG_FQ_R2_INIT {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, \
                        0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u}
#define CPG_FR_P_INIT {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}
#define CPG_FR_R_INIT {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}
#define CPG_FR_R2_INIT {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}

static const uint32_t H_FQ_P[12] = CPG_FQ_P_INIT;
static const uint32_t H_FQ_R[12] = CPG_FQ_R_INIT;
static const uint32_t H_FQ_R2[12] = CPG_FQ_R2_INIT;
static const uint32_t H_FR_P[8] = CPG_FR_P_INIT;
static const uint32_t H_FR_R[8] = CPG_FR_R_INIT;
static const uint32_t H_FR_R2[8] = CPG_FR_R2_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t D_FQ_P[12] = CPG_FQ_P_INIT;
static __device__ __constant__ uint32_t D_FQ_R[12] = CPG_FQ_R_INIT;
static __device__ __constant__ uint32_t D_FQ_R2[12] = CPG_FQ_R2_INIT;
static __device__ __constant__ uint32_t D_FR_P[8] = CPG_FR_P_INIT;
static __device__ __constant__ uint32_t D_FR_R[8] = CPG_FR_R_INIT;
static __device__ __constant__ uint32_t D_FR_R2[8] = CPG_FR_R2_INIT;
#endif

#ifdef __CUDA_ARCH__
#define CPG_SEL(name) D_##name
#else
#define CPG_SEL(name) H_##name
#endif

struct FqCfg {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    static constexpr bool FAST_SQR = true;        // p < 2^381: three spare bits, mont_sqr_n's partial sums fit
    static CPG_HD const uint32_t* p() { return CPG_SEL(FQ_P); }
    static CPG_HD const uint32_t* one() { return CPG_SEL(FQ_R); }
    static CPG_HD const uint32_t* r2() { return CPG_SEL(FQ_R2); }
};
struct FrCfg {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    static constexpr bool FAST_SQR = false;       // r > 2^254: the doubled operand overflows mont_sqr_n's window
    static CPG_HD const uint32_t* p() { return CPG_SEL(FR_P); }
    static CPG_HD const uint32_t* one() { return CPG_SEL(FR_R); }
    static CPG_HD const uint32_t* r2() { return CPG_SEL(FR_R2); }
};

template <class C>
struct Fp {
    static constexpr int N = C::N;
    uint32_t l[N];

    static CPG_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    static CPG_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = C::one()[i];
        return r;
    }
    CPG_HD bool is_zero() const { return is_zero_n<N>(l); }
    CPG_HD bool operator==(const Fp& o) const { return eq_n<N>(l, o.l); }
    CPG_HD bool operator!=(const Fp& o) const { return !eq_n<N>(l, o.l); }
};

template <class C> CPG_HD Fp<C> mul(const Fp<C>& a, const Fp<C>& b) { Fp<C> r; mont_mul_n<C::N>(r.l, a.l, b.l, C::p(), C::INV); return r; }
template <class C> CPG_HD Fp<C> sqr(const Fp<C>& a) {
    Fp<C> r;
    if (C::FAST_SQR) mont_sqr_n<C::N>(r.l, a.l, C::p(), C::INV);
    else mont_mul_n<C::N>(r.l, a.l, a.l, C::p(), C::INV);
    return r;
}
// a^(2^n).  The device build of Fq overrides it with ONE register-ABI call that loops inside the callee.
template <class C> CPG_HD Fp<C> sqr_n(Fp<C> a, int n) { for (int i = 0; i < n; i++) a = sqr(a); return a; }
template <class C> CPG_HD Fp<C> add(const Fp<C>& a, const Fp<C>& b) { Fp<C> r; mod_add_n<C::N>(r.l, a.l, b.l, C::p()); return r; }
template <class C> CPG_HD Fp<C> sub(const Fp<C>& a, const Fp<C>& b) { Fp<C> r; mod_sub_n<C::N>(r.l, a.l, b.l, C::p()); return r; }
template <class C> CPG_HD Fp<C> dbl(const Fp<C>& a) { return add(a, a); }
template <class C> CPG_HD Fp<C> neg(const Fp<C>& a) { return sub(Fp<C>::zero(), a); }
// plain integer (< p) -> Montgomery form and back
template <class C> CPG_HD Fp<C> to_mont(const Fp<C>& a) { Fp<C> r2; for (int i = 0; i < C::N; i++) r2.l[i] = C::r2()[i]; return mul(a, r2); }
template <class C> CPG_HD Fp<C> from_mont(const Fp<C>& a) { Fp<C> o = Fp<C>::zero(); o.l[0] = 1; return mul(a, o); }
// true iff the plain integer a is < p
template <class C> CPG_HD bool is_canonical(const Fp<C>& a) { return !geq_n<C::N>(a.l, C::p()); }

// Lazy forms for exponentiation chains of Fq (values in [0, 2p), bigint.cuh mont_mul_n<N, true>): LZ = false is the
// plain product.  Only for fields with three spare bits (FAST_SQR).
template <bool LZ, class C> CPG_HD Fp<C> mul_lz(const Fp<C>& a, const Fp<C>& b) {
    if constexpr (!LZ) return mul(a, b);
    else { Fp<C> r; mont_mul_n<C::N, true>(r.l, a.l, b.l, C::p(), C::INV); return r; }
}
template <bool LZ, class C> CPG_HD Fp<C> sqr_n_lz(Fp<C> a, int n) {
    if constexpr (!LZ) return sqr_n(a, n);
    else {
        for (int i = 0; i < n; i++) mont_sqr_n<C::N, true>(a.l, a.l, C::p(), C::INV);
        return a;
    }
}
template <class C> CPG_HD Fp<C> reduce_once(const Fp<C>& a) { Fp<C> r; reduce_once_n<C::N>(r.l, a.l, C::p()); return r; }

// a^e for a public exponent given as little-endian u32 words: left-to-right sliding window of 5 bits over
// the 16 odd powers a, a^3 .. a^31.  For the 379-bit sqrt / inversion exponents of Fq that is 376 squarings +
// 81 products (a fixed 4-bit window needs 105 products).  The exponent is uniform across the warp, so
// neither the window scan nor the table index diverges.
// LZ = true (Fq only): every intermediate stays in [0, 2p) without its conditional subtraction (24 + 12 of the ~380
// instructions of a product / squaring, which on B200 do NOT hide under the IMAD.WIDE stream - profiles/r02_pipe_probe.txt),
// one reduction at the end.
template <class C, int EW, bool LZ = false>
CPG_HD Fp<C> pow_public(const Fp<C>& a, const uint32_t (&e)[EW]) {
    static_assert(!LZ || C::FAST_SQR, "lazy chains need the field's three spare bits");
    Fp<C> tbl[16];
    tbl[0] = a;
    Fp<C> a2 = sqr_n_lz<LZ>(a, 1);
    for (int i = 1; i < 16; i++) tbl[i] = mul_lz<LZ>(tbl[i - 1], a2);
    Fp<C> acc = Fp<C>::one();
    bool started = false;
    int i = EW * 32 - 1;
    while (i >= 0) {
        if (!((e[i >> 5] >> (i & 31)) & 1u)) {                // a run of clear bits: that many squarings in one call
            int z = 1;
            while (i - z >= 0 && !((e[(i - z) >> 5] >> ((i - z) & 31)) & 1u)) z++;
            if (started) acc = sqr_n_lz<LZ>(acc, z);
            i -= z;
            continue;
        }
        int l = i >= 4 ? 5 : i + 1;                         // longest window <= 5 bits that ends in a set bit
        while (!((e[(i - l + 1) >> 5] >> ((i - l + 1) & 31)) & 1u)) l--;
        uint32_t v = 0;
        for (int k = 0; k < l; k++) v = (v << 1) | ((e[(i - k) >> 5] >> ((i - k) & 31)) & 1u);
        if (started) {
            acc = mul_lz<LZ>(sqr_n_lz<LZ>(acc, l), tbl[v >> 1]);
        } else {
            acc = tbl[v >> 1];
            started = true;
        }
        i -= l;
    }
    if constexpr (LZ) return reduce_once(acc);
    else return acc;
}

typedef Fp<FqCfg> Fq;
typedef Fp<FrCfg> Fr;

// Fq products as real calls (device only, -DCPG_FIELD_CALLS): the ABI passes both operands and the
// result in registers (no stack traffic), each call costs ~50 register moves, and the kernels shrink
// from ~80 KB of straight-line SASS per point addition to a ~5 KB multiply body that stays in the
// instruction cache.  Measured choice: see DESIGN.md "instruction cache".
#if defined(__CUDA_ARCH__) && defined(CPG_FIELD_CALLS)
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { Fq r; mont_mul_n<12>(r.l, a.l, b.l, FqCfg::p(), FqCfg::INV); return r; }
static __device__ __noinline__ Fq fq_sqr_call(Fq a) { Fq r; mont_sqr_n<12>(r.l, a.l, FqCfg::p(), FqCfg::INV); return r; }
__device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
__device__ __forceinline__ Fq sqr(const Fq& a) { return fq_sqr_call(a); }
// repeated squaring inside ONE call: the ~50 register moves of the call ABI are paid once per run, not per squaring
static __device__ __noinline__ Fq fq_sqr_n_call(Fq a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) mont_sqr_n<12>(a.l, a.l, FqCfg::p(), FqCfg::INV);
    return a;
}
__device__ __forceinline__ Fq sqr_n(Fq a, int n) { return fq_sqr_n_call(a, n); }
static __device__ __noinline__ Fq fq_mul_lz_call(Fq a, Fq b) { Fq r; mont_mul_n<12, true>(r.l, a.l, b.l, FqCfg::p(), FqCfg::INV); return r; }
static __device__ __noinline__ Fq fq_sqr_n_lz_call(Fq a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) mont_sqr_n<12, true>(a.l, a.l, FqCfg::p(), FqCfg::INV);
    return a;
}
template <> __device__ __forceinline__ Fq mul_lz<true, FqCfg>(const Fq& a, const Fq& b) { return fq_mul_lz_call(a, b); }
template <> __device__ __forceinline__ Fq sqr_n_lz<true, FqCfg>(Fq a, int n) { return fq_sqr_n_lz_call(a, n); }
#endif

// Fq inversion, two ways.
// (1) Fermat, a^(p-2): 376 squarings + 83 products on the multiply pipe.  Kept as the cross-check of (2).
CPG_HD Fq fq_inv_fermat(const Fq& a) {
    const uint32_t e[12] = {0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                            0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return pow_public<FqCfg, 12>(a, e);
}
// (2) Bernstein-Yang "safegcd" division steps, the variant that keeps zeta = -(delta + 1/2), on signed 30-bit limbs:
// 30 branch-free division steps on the low words of (f, g) build a 2x2 transition matrix t with entries in
// [-2^30, 2^30], which is then applied to the full-size (f, g) and, modulo p, to (d, e):
//     [f, g] <- t [f, g] / 2^30        [d, e] <- t [d, e] / 2^30  (mod p)
// invariant d x = f, e x = g (mod p); when g reaches 0, f = +-1 and d = +-x^-1.  About 30 rounds for a 381-bit modulus:
// ~4 000 wide multiplies + ~16 000 ALU instructions instead of the ~133 000 wide multiplies of (1).  The loop ends when
// g = 0, so correctness does not rest on an iteration bound.  inverse(0) = 0, as in (1).
struct S30 { int32_t v[13]; };                      // value = sum v[i] 2^(30 i), limbs in (-2^30, 2^30), top limb carries the sign
#define CPG_FQ_P30_INIT {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, \
                         0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111}
#define CPG_FQ_R3_INIT {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, \
                        0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u}
static const int32_t H_FQ_P30[13] = CPG_FQ_P30_INIT;
static const uint32_t H_FQ_R3[12] = CPG_FQ_R3_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ int32_t D_FQ_P30[13] = CPG_FQ_P30_INIT;
static __device__ __constant__ uint32_t D_FQ_R3[12] = CPG_FQ_R3_INIT;
#endif
constexpr int32_t FQ_M30 = 0x3fffffff;
constexpr uint32_t FQ_PINV30 = 0x00030003u;         // p^-1 mod 2^30

struct Trans30 { int32_t u, v, q, r; };
// 30 division steps on the low words; returns the new zeta
CPG_HD int32_t fq_divsteps30(int32_t zeta, uint32_t f, uint32_t g, Trans30& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;             // signed values mod 2^32 (left shifts stay defined)
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        uint32_t m1 = (uint32_t)(zeta >> 31);        // zeta < 0
        const uint32_t m2 = 0u - (g & 1u);           // g odd
        const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;   // f, u, v negated if zeta < 0
        g += x & m2; q += y & m2; r += z & m2;
        m1 &= m2;
        zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;   // zeta < 0 and g odd: -zeta - 2, else zeta - 1
        f += g & m1; u += q & m1; v += r & m1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}
// [d, e] <- t [d, e] / 2^30 mod p, for d, e in (-2p, p)
CPG_HD void fq_update_de30(S30& d, S30& e, const Trans30& t) {
    const int32_t* P30 = CPG_SEL(FQ_P30);
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[12] >> 31, se = e.v[12] >> 31;
    int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);      // multiples of p that undo a negative d / e
    int32_t di = d.v[0], ei = e.v[0];
    int64_t cd = (int64_t)u * di + (int64_t)v * ei, ce = (int64_t)q * di + (int64_t)r * ei;
    md -= (int32_t)((FQ_PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)FQ_M30);   // make the low 30 bits of t [d, e] + p [md, me] vanish
    me -= (int32_t)((FQ_PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)FQ_M30);
    cd += (int64_t)P30[0] * md; ce += (int64_t)P30[0] * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 13; i++) {
        di = d.v[i]; ei = e.v[i];
        cd += (int64_t)u * di + (int64_t)v * ei + (int64_t)P30[i] * md;
        ce += (int64_t)q * di + (int64_t)r * ei + (int64_t)P30[i] * me;
        d.v[i - 1] = (int32_t)cd & FQ_M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & FQ_M30; ce >>= 30;
    }
    d.v[12] = (int32_t)cd; e.v[12] = (int32_t)ce;
}
// [f, g] <- t [f, g] / 2^30 (exact)
CPG_HD void fq_update_fg30(S30& f, S30& g, const Trans30& t) {
    const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t fi = f.v[0], gi = g.v[0];
    int64_t cf = (int64_t)u * fi + (int64_t)v * gi, cg = (int64_t)q * fi + (int64_t)r * gi;
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < 13; i++) {
        fi = f.v[i]; gi = g.v[i];
        cf += (int64_t)u * fi + (int64_t)v * gi;
        cg += (int64_t)q * fi + (int64_t)r * gi;
        f.v[i - 1] = (int32_t)cf & FQ_M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & FQ_M30; cg >>= 30;
    }
    f.v[12] = (int32_t)cf; g.v[12] = (int32_t)cg;
}
// r in (-2p, p) -> sign * r in [0, p)   (sign < 0 negates)
CPG_HD void fq_normalize30(S30& r, int32_t sign) {
    const int32_t* P30 = CPG_SEL(FQ_P30);
    int32_t add = r.v[12] >> 31;
    const int32_t ng = sign >> 31;
#pragma unroll
    for (int i = 0; i < 13; i++) { int32_t x = r.v[i] + (P30[i] & add); r.v[i] = (x ^ ng) - ng; }
#pragma unroll
    for (int i = 0; i < 12; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= FQ_M30; }
    add = r.v[12] >> 31;
#pragma unroll
    for (int i = 0; i < 13; i++) r.v[i] += P30[i] & add;
#pragma unroll
    for (int i = 0; i < 12; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= FQ_M30; }
}
CPG_HD Fq fq_inv_safegcd(const Fq& a) {               // a and the result in Montgomery form
    S30 d, e, f, g;
    // 12 x 32-bit words -> 13 x 30-bit limbs
#pragma unroll
    for (int i = 0; i < 13; i++) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t x = a.l[w] >> sh;
        if (sh > 2 && w + 1 < 12) x |= a.l[w + 1] << (32 - sh);
        g.v[i] = (int32_t)(x & (uint32_t)FQ_M30);
        f.v[i] = CPG_SEL(FQ_P30)[i];
        d.v[i] = 0; e.v[i] = 0;
    }
    e.v[0] = 1;
    int32_t zeta = -1;
    for (int it = 0; it < 64; it++) {
        Trans30 t;
        zeta = fq_divsteps30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        fq_update_de30(d, e, t);
        fq_update_fg30(f, g, t);
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 13; i++) nz |= g.v[i];
        // the lanes of a warp leave together (a lane whose g is already 0 idles through harmless extra rounds:
        // t = diag(2^30, 1) leaves f and d as they are) - lanes that drift apart here rarely rejoin
#ifdef __CUDA_ARCH__
        if (!__any_sync(__activemask(), nz != 0)) break;
#else
        if (nz == 0) break;
#endif
    }
    fq_normalize30(d, f.v[12]);                         // d = +-(a R)^-1 with the sign of f = +-1
    Fq x;                                               // back to 12 x 32-bit words
#pragma unroll
    for (int w = 0; w < 12; w++) {
        const int bit = 32 * w, i = bit / 30, sh = bit % 30;
        uint32_t lo = (uint32_t)d.v[i] >> sh;
        uint32_t acc = lo | ((uint32_t)d.v[i + 1] << (30 - sh));
        if (30 - sh + 30 < 32 && i + 2 < 13) acc |= (uint32_t)d.v[i + 2] << (60 - sh);
        x.l[w] = acc;
    }
    Fq r3;
#pragma unroll
    for (int i = 0; i < 12; i++) r3.l[i] = CPG_SEL(FQ_R3)[i];
    return mul(x, r3);                                  // (a R)^-1 R^3 R^-1 = a^-1 R
}
// one copy of the inversion per kernel (register ABI, like the products above) instead of one per call site
#if defined(__CUDA_ARCH__) && defined(CPG_FIELD_CALLS)
static __device__ __noinline__ Fq fq_inv_call(Fq a) { return fq_inv_safegcd(a); }
__device__ __forceinline__ Fq fq_inv(const Fq& a) { return fq_inv_call(a); }
#else
CPG_HD Fq fq_inv(const Fq& a) { return fq_inv_safegcd(a); }
#endif
CPG_HD Fq fq_sqrt_candidate(const Fq& a) {
    const uint32_t e[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                            0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};
    return pow_public<FqCfg, 12, true>(a, e);
}
CPG_HD Fr fr_inv(const Fr& a) {
    const uint32_t e[8] = {0xffffffffu, 0xfffffffeu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return pow_public<FrCfg, 8>(a, e);
}

// (p-1)/2 as a plain integer: y is "lexicographically largest" iff y > (p-1)/2.
CPG_HD bool fq_is_lex_largest(const Fq& y_plain) {
    const uint32_t half[12] = {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u,
                               0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u};
    // y > half  <=>  !(half >= y)
    return !geq_n<12>(half, y_plain.l);
}

}  // namespace cpg

// Multi-limb (32-bit) carry-chain primitives and Montgomery arithmetic, sm_100a.
//
// On the device every primitive is ONE PTX instruction using the carry flag
// (add.cc / addc.cc / mad.lo.cc / madc.hi.cc ...).  ptxas fuses each
// {mad.lo.cc, madc.hi.cc} pair on the same operands into a single IMAD.WIDE.U32(.X),
// so a 12-limb Montgomery product is 2*12^2 + 12 = 300 integer-pipe MACs.
//
// On the host (no __CUDA_ARCH__) the same primitives are emulated with an explicit
// carry variable so the *identical* algorithm code is unit-tested on the CPU
// (tests/test_host_field.py); that emulation is a test seam, not a product path.
#pragma once
#include <stdint.h>

// Forced inlining is for the DEVICE pass only.  nvcc's host pass compiles every __host__ __device__ functor as well;
// with always_inline there gcc flattens whole point-arithmetic call trees into each of them (5.5 of the 7.5 minutes a
// build of this library took), for code the host never runs hot.
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define CPG_HD __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define CPG_HD __host__ __device__ inline
#else
#define CPG_HD inline
#endif

namespace cpg {

#ifdef __CUDA_ARCH__
CPG_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
CPG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CPG_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
CPG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
CPG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
CPG_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
CPG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
CPG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
CPG_HD uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// ---- host emulation of the PTX carry flag (test seam) ----
static thread_local uint32_t emu_cf = 0;
CPG_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + emu_cf; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + emu_cf; }
CPG_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; emu_cf = (uint32_t)(d >> 63); return (uint32_t)d; }
CPG_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - emu_cf; emu_cf = (uint32_t)(d >> 63); return (uint32_t)d; }
CPG_HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - emu_cf; }
CPG_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CPG_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
CPG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(a * b) + c; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(a * b) + c + emu_cf; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c + emu_cf; emu_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
CPG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + emu_cf; }
CPG_HD uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c + emu_cf; }
#endif

// acc[0..N) += (a[0] + a[2]*2^64 + a[4]*2^128 ...) * b ; the carry out of limb N-1 is left in CF.
template <int N>
CPG_HD void cmad_even(uint32_t* acc, const uint32_t* a, uint32_t b) {
    acc[0] = mad_lo_cc(a[0], b, acc[0]);
    acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(a[j], b, acc[j]);
        acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 1]);
    }
}

// x[j] = x[j+2] + (a[0] + a[2]*2^64 ...)*b  (x[j >= N] read as 0), carry-in from CF, in place.
template <int N>
CPG_HD void madc_shift2(uint32_t* x, const uint32_t* a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        x[j] = madc_lo_cc(a[j], b, x[j + 2]);
        x[j + 1] = madc_hi_cc(a[j], b, x[j + 3]);
    }
    x[N - 2] = madc_lo_cc(a[N - 2], b, 0);
    x[N - 1] = madc_hi(a[N - 2], b, 0);
}

// One interleaved Montgomery step.  X is aligned at limb 0, Y at limb 1 (value = X + 2^32 * Y).
// On entry Yold (aligned 0 in the previous frame, limb 0 already zero) is passed as `y` and the
// previous Y as `x`; dividing by 2^32 swaps their roles.  See DESIGN.md "Fq multiplication".
template <int N>
CPG_HD void mont_step(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t bi, const uint32_t* p, uint32_t inv) {
    x[0] = add_cc(x[0], y[1]);          // stray limb of the old aligned-0 accumulator
    madc_shift2<N>(y, a + 1, bi);       // y <- (y >> 64) + a_odd * bi   (+ carry of the line above)
    cmad_even<N>(x, a, bi);             // x += a_even * bi
    y[N - 1] = addc(y[N - 1], 0);
    uint32_t m = mul_lo(x[0], inv);
    cmad_even<N>(x, p, m);              // x += p_even * m  (limb 0 becomes 0)
    y[N - 1] = addc(y[N - 1], 0);
    cmad_even<N>(y, p + 1, m);          // y += p_odd * m
}

// 32 x 32 -> 64 as ONE IMAD.WIDE.U32 (a separate a*b and __umulhi(a, b) compile to IMAD + IMAD.HI)
CPG_HD void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
}

// r = a * b / 2^(32N) mod p, all operands < p, result < p.  r may alias a or b.
// LAZY = true leaves out the final conditional subtraction: for p < 2^(32N-3) (Fq) operands < 2p give
// t = (ab + mp) / R < p (4p/R + 1) < 1.41 p and every interleaved partial sum stays < 3p < 2^(32N), so a chain of
// products / squarings can run on values in [0, 2p) and reduce ONCE at its end (reduce_once_n).
template <int N, bool LAZY = false>
CPG_HD void mont_mul_n(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv) {
    uint32_t e[N], o[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        mul_wide(a[j], b[0], e[j], e[j + 1]);
        mul_wide(a[j + 1], b[0], o[j], o[j + 1]);
    }
    {
        uint32_t m = mul_lo(e[0], inv);
        cmad_even<N>(e, p, m);
        o[N - 1] = addc(o[N - 1], 0);
        cmad_even<N>(o, p + 1, m);
    }
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_step<N>(o, e, a, b[i], p, inv);
        if (i + 1 < N) mont_step<N>(e, o, a, b[i + 1], p, inv);
    }
    // N even: the aligned-0 accumulator is `o` (limb 0 == 0), the aligned-1 one is `e`.
    uint32_t t[N];
    t[0] = add_cc(o[1], e[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) t[j] = addc_cc(o[j + 1], e[j]);
    t[N - 1] = addc(e[N - 1], 0);
    if (LAZY) {
#pragma unroll
        for (int j = 0; j < N; j++) r[j] = t[j];
        return;
    }
    // t < 2p: one conditional subtraction
    uint32_t s[N];
    s[0] = sub_cc(t[0], p[0]);
#pragma unroll
    for (int j = 1; j < N; j++) s[j] = subc_cc(t[j], p[j]);
    uint32_t borrow = subc(0, 0);  // 0xffffffff if t < p
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = borrow ? t[j] : s[j];
}

// ---- dedicated squaring ------------------------------------------------------------------------
// a^2 = sum_i a_i * f^(i) * 2^(32 i)  with the row vector  f^(i) = [0,...,0, a_i, 2a_{i+1}, 2a_{i+2}, ...]
// (i leading zeros): the 66 off-diagonal products are taken once and doubled through the operand
// d = 2a, the 12 diagonal ones once - 78 instead of 144 product MACs; the 144 reduction MACs stay.
// Needs p < 2^(32N-3): d = 2a is then exact in N limbs and the look-ahead partial sums
// (row i already holds 2 a_i a_j for all j > i) stay below 2^(32N).  True for Fq (381 bits in 384), not for Fr.  Row i of the interleaved (CIOS, even/odd) scheme of
// mont_mul_n then uses f^(i) as its multiplicand vector and a_i as its multiplier; entries below i are
// compile-time zeros whose MACs degenerate to the carry/shift adds.
template <int N, int I>
CPG_HD void sqr_row_vec(uint32_t* f, const uint32_t* a, const uint32_t* d) {
#pragma unroll
    for (int j = 0; j < N; j++) f[j] = j < I ? 0u : (j == I ? a[j] : (j == I + 1 ? (d[j] & 0xfffffffeu) : d[j]));
}
// x += (f[R] + f[R+2] 2^64 + ...) * b over even entries R >= I only; carry out left in CF.
// Returns false (and leaves CF untouched) when the row has no even entry >= I.
template <int N, int I>
CPG_HD bool cmad_even_from(uint32_t* acc, const uint32_t* f, uint32_t b) {
    constexpr int R0 = (I + 1) & ~1;      // first even index >= I
    if (R0 >= N) return false;
    acc[R0] = mad_lo_cc(f[R0], b, acc[R0]);
    acc[R0 + 1] = madc_hi_cc(f[R0], b, acc[R0 + 1]);
#pragma unroll
    for (int j = R0 + 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(f[j], b, acc[j]);
        acc[j + 1] = madc_hi_cc(f[j], b, acc[j + 1]);
    }
    return true;
}
// y[j] = y[j+2] + (odd entries f[j+1], j+1 >= I) * b, carry-in from CF, in place (cf. madc_shift2)
template <int N, int I>
CPG_HD void madc_shift2_from(uint32_t* y, const uint32_t* f, uint32_t b) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        if (j + 1 >= I) {
            y[j] = madc_lo_cc(f[j + 1], b, y[j + 2]);
            y[j + 1] = madc_hi_cc(f[j + 1], b, y[j + 3]);
        } else {
            y[j] = addc_cc(y[j + 2], 0);
            y[j + 1] = addc_cc(y[j + 3], 0);
        }
    }
    if (N - 1 >= I) {
        y[N - 2] = madc_lo_cc(f[N - 1], b, 0);
        y[N - 1] = madc_hi(f[N - 1], b, 0);
    } else {
        y[N - 2] = addc(0, 0);
        y[N - 1] = 0;
    }
}
template <int N, int I>
CPG_HD void mont_sqr_step(uint32_t* x, uint32_t* y, const uint32_t* a, const uint32_t* d, const uint32_t* p, uint32_t inv) {
    uint32_t f[N];
    sqr_row_vec<N, I>(f, a, d);
    const uint32_t bi = a[I];
    x[0] = add_cc(x[0], y[1]);
    madc_shift2_from<N, I>(y, f, bi);
    if (cmad_even_from<N, I>(x, f, bi)) y[N - 1] = addc(y[N - 1], 0);
    uint32_t m = mul_lo(x[0], inv);
    cmad_even<N>(x, p, m);
    y[N - 1] = addc(y[N - 1], 0);
    cmad_even<N>(y, p + 1, m);
}
template <int N, int I>
struct SqrSteps {
    static CPG_HD void run(uint32_t* e, uint32_t* o, const uint32_t* a, const uint32_t* d, const uint32_t* p, uint32_t inv) {
        // odd I: the accumulator aligned at 1 is `o` (see mont_mul_n), even I: `e`
        if (I & 1) mont_sqr_step<N, I>(o, e, a, d, p, inv);
        else mont_sqr_step<N, I>(e, o, a, d, p, inv);
        SqrSteps<N, I + 1>::run(e, o, a, d, p, inv);
    }
};
template <int N>
struct SqrSteps<N, N> {
    static CPG_HD void run(uint32_t*, uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t) {}
};
// r = a * a / 2^(32N) mod p, a < p < 2^(32N-3), result < p.  r may alias a.
// LAZY = true: a < 2p allowed (d = 2a < 4p < 2^(32N-1) is still exact, the look-ahead partial sums are < 2a + p < 5p
// < 2^(32N)), no final subtraction, result < 1.41 p.
template <int N, bool LAZY = false>
CPG_HD void mont_sqr_n(uint32_t* r, const uint32_t* a, const uint32_t* p, uint32_t inv) {
    uint32_t d[N];
    d[0] = a[0] << 1;
#pragma unroll
    for (int j = 1; j < N; j++) d[j] = (a[j] << 1) | (a[j - 1] >> 31);
    uint32_t e[N], o[N];
    {   // row 0: vector [a_0, 2a_1 (bit 0 clear), d_2, ...] times a_0
        uint32_t f[N];
        sqr_row_vec<N, 0>(f, a, d);
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            mul_wide(f[j], a[0], e[j], e[j + 1]);
            mul_wide(f[j + 1], a[0], o[j], o[j + 1]);
        }
        uint32_t m = mul_lo(e[0], inv);
        cmad_even<N>(e, p, m);
        o[N - 1] = addc(o[N - 1], 0);
        cmad_even<N>(o, p + 1, m);
    }
    SqrSteps<N, 1>::run(e, o, a, d, p, inv);
    uint32_t t[N];
    t[0] = add_cc(o[1], e[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) t[j] = addc_cc(o[j + 1], e[j]);
    t[N - 1] = addc(e[N - 1], 0);
    if (LAZY) {
#pragma unroll
        for (int j = 0; j < N; j++) r[j] = t[j];
        return;
    }
    uint32_t s[N];
    s[0] = sub_cc(t[0], p[0]);
#pragma unroll
    for (int j = 1; j < N; j++) s[j] = subc_cc(t[j], p[j]);
    uint32_t borrow = subc(0, 0);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = borrow ? t[j] : s[j];
}
// t in [0, 2p) -> [0, p)
template <int N>
CPG_HD void reduce_once_n(uint32_t* r, const uint32_t* t, const uint32_t* p) {
    uint32_t s[N];
    s[0] = sub_cc(t[0], p[0]);
#pragma unroll
    for (int j = 1; j < N; j++) s[j] = subc_cc(t[j], p[j]);
    uint32_t borrow = subc(0, 0);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = borrow ? t[j] : s[j];
}

// r = a + b mod p
template <int N>
CPG_HD void mod_add_n(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* p) {
    uint32_t t[N], s[N];
    t[0] = add_cc(a[0], b[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) t[j] = addc_cc(a[j], b[j]);
    t[N - 1] = addc(a[N - 1], b[N - 1]);  // 2p < 2^(32N): no carry out
    s[0] = sub_cc(t[0], p[0]);
#pragma unroll
    for (int j = 1; j < N; j++) s[j] = subc_cc(t[j], p[j]);
    uint32_t borrow = subc(0, 0);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = borrow ? t[j] : s[j];
}

// r = a - b mod p
template <int N>
CPG_HD void mod_sub_n(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* p) {
    uint32_t t[N];
    t[0] = sub_cc(a[0], b[0]);
#pragma unroll
    for (int j = 1; j < N; j++) t[j] = subc_cc(a[j], b[j]);
    uint32_t borrow = subc(0, 0);
    r[0] = add_cc(t[0], p[0] & borrow);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(t[j], p[j] & borrow);
    r[N - 1] = addc(t[N - 1], p[N - 1] & borrow);
}

template <int N>
CPG_HD bool is_zero_n(const uint32_t* a) {
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < N; j++) t |= a[j];
    return t == 0;
}
template <int N>
CPG_HD bool eq_n(const uint32_t* a, const uint32_t* b) {
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < N; j++) t |= a[j] ^ b[j];
    return t == 0;
}
// a >= b on plain integers
template <int N>
CPG_HD bool geq_n(const uint32_t* a, const uint32_t* b) {
    uint32_t t = sub_cc(a[0], b[0]);
#pragma unroll
    for (int j = 1; j < N; j++) t = subc_cc(a[j], b[j]);
    (void)t;
    return subc(0, 0) == 0;
}

}  // namespace cpg

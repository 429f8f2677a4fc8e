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

#if defined(__CUDACC__)
#define CPG_HD __host__ __device__ __forceinline__
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

// r = a * b / 2^(32N) mod p, all operands < p, result < p.  r may alias a or b.
template <int N>
CPG_HD void mont_mul_n(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv) {
    uint32_t e[N], o[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        e[j] = mul_lo(a[j], b[0]);
        e[j + 1] = mul_hi(a[j], b[0]);
        o[j] = mul_lo(a[j + 1], b[0]);
        o[j + 1] = mul_hi(a[j + 1], b[0]);
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
    // t < 2p: one conditional subtraction
    uint32_t s[N];
    s[0] = sub_cc(t[0], p[0]);
#pragma unroll
    for (int j = 1; j < N; j++) s[j] = subc_cc(t[j], p[j]);
    uint32_t borrow = subc(0, 0);  // 0xffffffff if t < p
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

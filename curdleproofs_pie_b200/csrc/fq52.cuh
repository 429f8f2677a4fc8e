// Fq on the FP64 pipe: 8 limbs of 52 bits held as integer-valued doubles, Montgomery radix 2^416.
//
// Why: a 12 x 32-bit Montgomery product is 300 IMAD.WIDE.U32 and B200 issues those at 32 lanes/clk/SM
// (profiles/r01_imad_probe.txt) - the multiply pipe is the roofline of every point kernel.  The FP64 pipe of the same
// SM (64 DFMA lanes/clk) is idle.  A 52 x 52-bit limb product is two DFMA and one DADD
//     hi = fma_rz(a, b, 2^104)                -> mantissa bits = floor(ab / 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)  -> mantissa bits = ab mod 2^52        (both exact)
// whose mantissas are accumulated per column as 64-bit integers on the ALU pipe (IADD3 takes two new terms per
// instruction pair; the exponent bits are cancelled by one precomputed constant per column).  A product is
// 64 + 64 limb products = 400 FP64-pipe instructions + ~290 ALU-pipe instructions and touches the multiply pipe only
// for the 8 quotient digits; a squaring takes 36 + 64 limb products.  Because the two field implementations sit on
// different pipes, warps running this one and warps running the integer one add up (see Decompress in msm.cuh).
//
// Values are kept in [0, p + eps): with operands < 2p the Montgomery result (ab + qp) / 2^416 is < p (1 + 2^-33),
// so no conditional subtraction is needed inside a chain; fq52_canon() does the single final one.
#pragma once
#include <stdint.h>

namespace cpg {

#define CPG_P52_INIT {0xeffffffffaaabull, 0xfeb153ffffb9full, 0x6b0f6241eabffull, 0x12bf6730d2a0full, \
                      0x764774b84f385ull, 0x1ba7b6434bacdull, 0x1ea397fe69a4bull, 0x1a011ull}
// 2^448 mod p: mm52(x 2^384, .) = x 2^416   (12 x u32 Montgomery form -> this form)
#define CPG_C384_416_INIT {0x7fde37dba9366ull, 0x4e27525bc342bull, 0x1f5b1e9778489ull, 0xb872b2b91b9dcull, \
                           0xb206f497dfcafull, 0x4137cc89a9b0bull, 0xd9d20d7e39959ull, 0x411cull}
// 2^384 mod p: mm52(x 2^416, .) = x 2^384   (back)
#define CPG_C416_384_INIT {0x900000002fffdull, 0xbc40c0002760ull, 0x3c758baebf400ull, 0x57455f4898575ull, \
                           0xd77ce58537052ull, 0x71a97a256ec6ull, 0xec3fa80e4935cull, 0x15f65ull}
// 2^416 mod p: the field's 1
#define CPG_R52_INIT {0x6480ea8e9b9afull, 0x65766c8fe444full, 0x8b540fea96f7dull, 0x3b2ee82efd422ull, \
                      0xa6723e5f0ade5ull, 0xff6eb6fdd4230ull, 0xe06ef23c24a25ull, 0x14c8eull}

#if defined(__CUDACC__)

static __device__ __constant__ uint64_t D_P52[8] = CPG_P52_INIT;
static __device__ __constant__ double D_P52D[8] = {(double)0xeffffffffaaabull, (double)0xfeb153ffffb9full, (double)0x6b0f6241eabffull, (double)0x12bf6730d2a0full,
                                                   (double)0x764774b84f385ull, (double)0x1ba7b6434bacdull, (double)0x1ea397fe69a4bull, (double)0x1a011ull};
static __device__ __constant__ uint64_t D_C384_416[8] = CPG_C384_416_INIT;
static __device__ __constant__ uint64_t D_C416_384[8] = CPG_C416_384_INIT;
static __device__ __constant__ uint64_t D_R52[8] = CPG_R52_INIT;

struct F52 { double v[8]; };

#define CPG52_MASK 0xfffffffffffffull
#define CPG52_NP 0x3fffcfffcfffdull              /* -p^-1 mod 2^52 */
#define CPG52_EXP_LO 0x4330000000000000ull       /* exponent bits of 2^52  + L */
#define CPG52_EXP_HI 0x4670000000000000ull       /* exponent bits of 2^104 + H 2^52 */

__device__ __forceinline__ double u52_to_double(uint64_t x) { return __longlong_as_double((long long)(x | CPG52_EXP_LO)) - 0x1p52; }

// number of (i, j) in [0,8)^2 with i + j == k
__host__ __device__ constexpr int cnt52(int k) { return (k < 0 || k > 14) ? 0 : (k < 8 ? k + 1 : 15 - k); }
// minus the exponent bits of every term column k will receive: T product-phase index sets + the reduction's
__host__ __device__ constexpr uint64_t col_bias52(int k, int lo_terms, int hi_terms) {
    return 0ull - ((uint64_t)lo_terms * CPG52_EXP_LO + (uint64_t)hi_terms * CPG52_EXP_HI);
}

__device__ __forceinline__ void limb_prod52(double a, double b, uint64_t& col, uint64_t& col1) {
    double hi = __fma_rz(a, b, 0x1p104);
    double lo = __fma_rz(a, b, (0x1p104 + 0x1p52) - hi);
    col += (uint64_t)__double_as_longlong(lo);
    col1 += (uint64_t)__double_as_longlong(hi);
}

// Montgomery reduction of the 16 columns c (already holding the product terms and the bias of all 128 + ... terms)
__device__ __forceinline__ void mont_reduce52(double (&r)[8], uint64_t (&c)[17]) {
    double p[8];
#pragma unroll
    for (int j = 0; j < 8; j++) p[j] = D_P52D[j];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t q = (c[i] * CPG52_NP) & CPG52_MASK;
        double qd = u52_to_double(q);
#pragma unroll
        for (int j = 0; j < 8; j++) limb_prod52(qd, p[j], c[i + j], c[i + j + 1]);
        c[i + 1] += c[i] >> 52;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t t = c[8 + j];
        c[9 + j] += t >> 52;
        r[j] = u52_to_double(t & CPG52_MASK);
    }
}

__device__ __forceinline__ void mm52(double (&r)[8], const double (&a)[8], const double (&b)[8]) {
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = col_bias52(k, 2 * cnt52(k), 2 * cnt52(k - 1));
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) limb_prod52(a[i], b[j], c[i + j], c[i + j + 1]);
    mont_reduce52(r, c);
}

// squaring: the 28 cross products once (doubled as integers), the 8 diagonal ones once
__device__ __forceinline__ void ms52(double (&r)[8], const double (&a)[8]) {
    uint64_t c[17];
    // cross terms (i < j): column k gets (cnt(k) - diag(k)) / 2 lo terms, same shifted for hi
#pragma unroll
    for (int k = 0; k < 17; k++) {
        int lo_x = (cnt52(k) - ((k % 2 == 0 && k <= 14) ? 1 : 0)) / 2;
        int hi_x = (cnt52(k - 1) - (((k - 1) % 2 == 0 && k - 1 >= 0 && k - 1 <= 14) ? 1 : 0)) / 2;
        c[k] = col_bias52(k, lo_x, hi_x);
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = i + 1; j < 8; j++) limb_prod52(a[i], a[j], c[i + j], c[i + j + 1]);
    // double, then the diagonal terms and the bias of diagonal + reduction terms
#pragma unroll
    for (int k = 0; k < 17; k++) {
        int lo_d = (k % 2 == 0 && k <= 14) ? 1 : 0;
        int hi_d = ((k - 1) % 2 == 0 && k - 1 >= 0 && k - 1 <= 14) ? 1 : 0;
        c[k] = c[k] + c[k] + col_bias52(k, lo_d + cnt52(k), hi_d + cnt52(k - 1));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) limb_prod52(a[i], a[i], c[2 * i], c[2 * i + 1]);
    mont_reduce52(r, c);
}

__device__ __forceinline__ F52 mul(const F52& a, const F52& b) { F52 r; mm52(r.v, a.v, b.v); return r; }
__device__ __forceinline__ F52 sqr(const F52& a) { F52 r; ms52(r.v, a.v); return r; }

// integer limbs (each < 2^52, value < 2^416) of a value, canonical: one conditional subtraction of p
__device__ __forceinline__ void fq52_canon(uint64_t (&o)[8], const F52& a) {
    uint64_t x[8], d[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = (uint64_t)__double_as_longlong(a.v[j] + 0x1p52) & CPG52_MASK;
    uint64_t borrow = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint64_t t = x[j] - D_P52[j] - borrow;
        borrow = t >> 63;
        d[j] = t & CPG52_MASK;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = borrow ? x[j] : d[j];
}

#endif  // __CUDACC__

}  // namespace cpg

// The sequential Fiat-Shamir transcript and the Fr bookkeeping that turns challenges into MSM
// coefficients.  Plain C++ in CPG_HD functions: the batched verifier runs this code either on host
// threads (the north star's default placement) or, one proof per thread, as a GPU kernel
//  - the same source either way.  Follows the reference, restated independently of oracle/:
//   Keccak-f[1600]      /root/reference/merlin_transcripts/merlin_transcripts/keccak.py:16-66
//   STROBE-128, R=166   merlin_transcripts/merlin_transcripts/strobe.py:16-107
//   Merlin framing      merlin_transcripts/merlin_transcripts/merlin_transcript.py:6-24
//   scalar challenges   curdleproofs/curdleproofs/curdleproofs_transcript.py:15-25
// KATs: tests/test_host_transcript.py runs merlin_transcripts/test_merlin.py:18,29,40's vectors through THIS code on the
// host and in a kernel (cpg_merlin_script, verify.inl).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef CPG_HD
#define CPG_HD inline
#endif
#if defined(__CUDACC__)
#define CPGH_CONST static __device__ __constant__
#else
#define CPGH_CONST static const
#endif
#ifdef __CUDA_ARCH__
#define CPGH_SEL(name) D_##name
#else
#define CPGH_SEL(name) H_##name
#endif

// The per-proof kernels (transcript + Fr algebra, one thread per proof) are latency-bound.  The Keccak permutation
// and the Fr inversion are real functions (one copy each instead of one per call site).  The Fr product stays inline:
// as a call its HFr operands travel through the stack (10 k local loads/stores in an 18 k-instruction VerifyPhase2),
// which is neutral for batches but 4x slower for a lone large proof; the STROBE absorb loop stays inline because as a
// function its byte accesses lose their address space (generic LD/ST instead of LDL/LDG).
#if defined(__CUDACC__)
#define CPGH_CALL __host__ __device__ __noinline__
#else
#define CPGH_CALL inline
#endif

namespace cpgh {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- Keccak-f[1600] ---
CPG_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
#define CPGH_KECCAK_RC_INIT { \
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL, \
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL, \
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, \
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL}
static const uint64_t H_KECCAK_RC[24] = CPGH_KECCAK_RC_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t D_KECCAK_RC[24] = CPGH_KECCAK_RC_INIT;
#endif

CPGH_CALL void keccak_f1600(uint64_t* A) {
    for (int r = 0; r < 24; r++) {
        uint64_t C0 = A[0] ^ A[5] ^ A[10] ^ A[15] ^ A[20], C1 = A[1] ^ A[6] ^ A[11] ^ A[16] ^ A[21];
        uint64_t C2 = A[2] ^ A[7] ^ A[12] ^ A[17] ^ A[22], C3 = A[3] ^ A[8] ^ A[13] ^ A[18] ^ A[23];
        uint64_t C4 = A[4] ^ A[9] ^ A[14] ^ A[19] ^ A[24];
        uint64_t D0 = C4 ^ rotl64(C1, 1), D1 = C0 ^ rotl64(C2, 1), D2 = C1 ^ rotl64(C3, 1), D3 = C2 ^ rotl64(C4, 1), D4 = C3 ^ rotl64(C0, 1);
        uint64_t B[25];
        // theta + rho + pi:  B[y + 5*((2x+3y)%5)] = rot(A[x+5y] ^ D[x], r[x,y])
        B[0] = A[0] ^ D0;
        B[10] = rotl64(A[1] ^ D1, 1);   B[20] = rotl64(A[2] ^ D2, 62);  B[5] = rotl64(A[3] ^ D3, 28);   B[15] = rotl64(A[4] ^ D4, 27);
        B[16] = rotl64(A[5] ^ D0, 36);  B[1] = rotl64(A[6] ^ D1, 44);   B[11] = rotl64(A[7] ^ D2, 6);   B[21] = rotl64(A[8] ^ D3, 55);  B[6] = rotl64(A[9] ^ D4, 20);
        B[7] = rotl64(A[10] ^ D0, 3);   B[17] = rotl64(A[11] ^ D1, 10); B[2] = rotl64(A[12] ^ D2, 43);  B[12] = rotl64(A[13] ^ D3, 25); B[22] = rotl64(A[14] ^ D4, 39);
        B[23] = rotl64(A[15] ^ D0, 41); B[8] = rotl64(A[16] ^ D1, 45);  B[18] = rotl64(A[17] ^ D2, 15); B[3] = rotl64(A[18] ^ D3, 21);  B[13] = rotl64(A[19] ^ D4, 8);
        B[14] = rotl64(A[20] ^ D0, 18); B[24] = rotl64(A[21] ^ D1, 2);  B[9] = rotl64(A[22] ^ D2, 61);  B[19] = rotl64(A[23] ^ D3, 56);  B[4] = rotl64(A[24] ^ D4, 14);
        for (int y = 0; y < 25; y += 5) {
            A[y + 0] = B[y + 0] ^ (~B[y + 1] & B[y + 2]);
            A[y + 1] = B[y + 1] ^ (~B[y + 2] & B[y + 3]);
            A[y + 2] = B[y + 2] ^ (~B[y + 3] & B[y + 4]);
            A[y + 3] = B[y + 3] ^ (~B[y + 4] & B[y + 0]);
            A[y + 4] = B[y + 4] ^ (~B[y + 0] & B[y + 1]);
        }
        A[0] ^= CPGH_SEL(KECCAK_RC)[r];
    }
}


// The same permutation by a WHOLE WARP for one state (device only).  Used when one proof is given a warp instead of a
// thread: all 32 lanes then run the transcript code in lock-step on identical private copies of the state, and only this
// function cooperates - lane l < 25 owns word l = x + 5y, a round is 18 shuffles (theta: 4 column-mates + 2 neighbour
// parities, pi: 1, chi: 2, each 64-bit = 2 x 32-bit) instead of ~150 dependent 64-bit operations on one thread, and at the
// end every lane collects the 25 words again.  ~6x shorter than keccak_f1600 on a lone thread; 32x its issue slots - the
// right trade when a batch is too small to hide the one-thread-per-proof latency (verify.inl: transcript mode).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
static __device__ __noinline__ void keccak_f1600_warp(uint64_t* A) {
    const int lane = (int)(threadIdx.x & 31u);
    const int l = lane < 25 ? lane : lane - 25;                 // lanes 25..31 shadow lanes 0..6 (results unused)
    const int x = l % 5, y = l / 5;
    // rho offsets r[x + 5y]
    const uint8_t RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    const int rot = RHO[l];
    const int col1 = (l + 5) % 25, col2 = (l + 10) % 25, col3 = (l + 15) % 25, col4 = (l + 20) % 25;
    const int xm1 = (x + 4) % 5 + 5 * y, xp1 = (x + 1) % 5 + 5 * y, xp2 = (x + 2) % 5 + 5 * y;
    const int pi_src = (x + 3 * y) % 5 + 5 * x;                 // B[x, y] = rot(A[(x + 3y) % 5, x])
    uint64_t a = A[l];
    for (int r = 0; r < 24; r++) {
        const uint64_t c = a ^ shfl64(a, col1) ^ shfl64(a, col2) ^ shfl64(a, col3) ^ shfl64(a, col4);   // parity of column x (in every row)
        const uint64_t cp = shfl64(c, xp1);
        a ^= shfl64(c, xm1) ^ ((cp << 1) | (cp >> 63));
        const uint64_t rt = rot ? ((a << rot) | (a >> (64 - rot))) : a;
        const uint64_t b = shfl64(rt, pi_src);
        a = b ^ (~shfl64(b, xp1) & shfl64(b, xp2));
        if (l == 0) a ^= D_KECCAK_RC[r];
    }
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = shfl64(a, i);
}
#endif

// --------------------------------------------------------------------- STROBE-128 ---
struct Strobe128 {
    static const int RATE = 166;
    enum { F_I = 1, F_A = 2, F_C = 4, F_T = 8, F_M = 16, F_K = 32 };
    union { uint64_t w[25]; uint8_t b[200]; } st;   // little-endian host assumed (x86-64 / aarch64)
    uint8_t pos, pos_begin, flags;
    uint8_t warp = 0;             // device: 1 = this state is driven by a whole warp in lock-step (keccak_f1600_warp)
    uint64_t permutations;
    CPG_HD void permute() {
#if defined(__CUDA_ARCH__)
        if (warp) { keccak_f1600_warp(st.w); return; }
#endif
        keccak_f1600(st.w);
    }

    CPG_HD void init(const uint8_t* label, size_t n) {
        memset(st.b, 0, 200);
        const uint8_t hdr[6] = {1, RATE + 2, 1, 0, 1, 96};
        memcpy(st.b, hdr, 6);
        const uint8_t ver[12] = {'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
        memcpy(st.b + 6, ver, 12);
        permute();
        pos = pos_begin = flags = 0;
        permutations = 1;
        meta_ad(label, n, false);
    }
    CPG_HD void run_f() {
        st.b[pos] ^= pos_begin;
        st.b[pos + 1] ^= 0x04;
        st.b[RATE + 1] ^= 0x80;
        permute();
        permutations++;
        pos = pos_begin = 0;
    }
    CPG_HD void absorb(const uint8_t* d, size_t n) {
        while (n) {
            size_t room = RATE - pos, take = n < room ? n : room;
            size_t i = 0;
            // 8 source bytes at a time when the source is 8-byte aligned (wire points and scalars are): the word is
            // XORed into the one or two state words it straddles.  Byte by byte this loop was 17 % of the instructions
            // of a verifier transcript (28 KB absorbed per proof, ~10 instructions per byte in device code).
            if ((((uintptr_t)d) & 7) == 0) {
                for (; i + 8 <= take; i += 8) {
                    uint64_t v;
#if defined(__CUDA_ARCH__)
                    v = *(const uint64_t*)(d + i);
#else
                    memcpy(&v, d + i, 8);
#endif
                    const size_t p = pos + i;
                    const unsigned sh = (unsigned)(p & 7) * 8;
                    st.w[p >> 3] ^= v << sh;
                    if (sh) st.w[(p >> 3) + 1] ^= v >> (64 - sh);
                }
            }
            for (; i < take; i++) st.b[pos + i] ^= d[i];
            pos += (uint8_t)take; d += take; n -= take;
            if (pos == RATE) run_f();
        }
    }
    CPG_HD void begin_op(uint8_t fl, bool more) {
        if (more) return;                         // continuation of the same operation
        uint8_t old = pos_begin;
        pos_begin = pos + 1;
        flags = fl;
        uint8_t hdr[2] = {old, fl};
        absorb(hdr, 2);
        if ((fl & (F_C | F_K)) && pos != 0) run_f();
    }
    CPG_HD void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(F_M | F_A, more); absorb(d, n); }
    CPG_HD void ad(const uint8_t* d, size_t n, bool more) { begin_op(F_A, more); absorb(d, n); }
    CPG_HD void key(const uint8_t* d, size_t n, bool more) {      // strobe.py:83-85 (not used by Merlin; kept for the STROBE KAT)
        begin_op(F_A | F_C, more);
        for (size_t i = 0; i < n; i++) {
            st.b[pos] = d[i];
            pos++;
            if (pos == RATE) run_f();
        }
    }
    CPG_HD void prf(uint8_t* out, size_t n, bool more) {
        begin_op(F_I | F_A | F_C, more);
        for (size_t i = 0; i < n; i++) {
            out[i] = st.b[pos];
            st.b[pos] = 0;
            pos++;
            if (pos == RATE) run_f();
        }
    }
};

// ------------------------------------------------------------------------- Fr (host) ---
// 4 x u64 Montgomery arithmetic mod r; values are kept in Montgomery form inside `HFr` (host Fr).
struct HFr { uint64_t l[4]; };
// 32-byte scalars / coefficient rows live at 8-byte-aligned offsets of every buffer of this library (rows start at
// multiples of 32, wire scalars at multiples of 16).  A plain memcpy / memset on a uint8_t* compiles to BYTE loads and
// stores in device code (40 k byte stores per proof and round when zeroing the coefficient rows), so 32-byte items move
// as four u64 whenever the pointer allows it, and rows are zeroed 8 bytes at a time.
CPG_HD void copy32(void* dst, const void* src) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 7) == 0) {
        uint64_t* d = (uint64_t*)dst; const uint64_t* s = (const uint64_t*)src;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    } else {
        memcpy(dst, src, 32);
    }
}
CPG_HD void zero_bytes(uint8_t* p, size_t n) {
    if ((((uintptr_t)p | n) & 7) == 0) {
        uint64_t* q = (uint64_t*)p;
        for (size_t i = 0; i < n / 8; i++) q[i] = 0;
    } else {
        memset(p, 0, n);
    }
}
#define CPGH_FR_MOD_INIT {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}
#define CPGH_FR_R1_INIT {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}  /* 2^256 mod r */
#define CPGH_FR_R2_INIT {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}  /* 2^512 mod r */
static const uint64_t H_FR_MOD[4] = CPGH_FR_MOD_INIT;
static const uint64_t H_FR_R1[4] = CPGH_FR_R1_INIT;
static const uint64_t H_FR_R2[4] = CPGH_FR_R2_INIT;
#define CPGH_FR_R3_INIT {0xc62c1807439b73afULL, 0x1b3e0d188cf06990ULL, 0x73d13c71c7b5f418ULL, 0x6e2a5bb9c8db33e9ULL}  /* 2^768 mod r */
static const uint64_t H_FR_R3[4] = CPGH_FR_R3_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t D_FR_R3[4] = CPGH_FR_R3_INIT;
static __device__ __constant__ uint64_t D_FR_MOD[4] = CPGH_FR_MOD_INIT;
static __device__ __constant__ uint64_t D_FR_R1[4] = CPGH_FR_R1_INIT;
static __device__ __constant__ uint64_t D_FR_R2[4] = CPGH_FR_R2_INIT;
#endif
#define FR_MOD CPGH_SEL(FR_MOD)
static const uint64_t FR_INV = 0xfffffffeffffffffULL;   // -r^-1 mod 2^64 (scalar constant: usable on both sides)
CPG_HD HFr fr_r1() { HFr r; for (int i = 0; i < 4; i++) r.l[i] = CPGH_SEL(FR_R1)[i]; return r; }
CPG_HD HFr fr_r2() { HFr r; for (int i = 0; i < 4; i++) r.l[i] = CPGH_SEL(FR_R2)[i]; return r; }

CPG_HD bool fr_geq_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) { if (a[i] > FR_MOD[i]) return true; if (a[i] < FR_MOD[i]) return false; }
    return true;
}
CPG_HD void fr_sub_mod(uint64_t* a) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - FR_MOD[i] - (uint64_t)br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
CPG_HD HFr fr_add(const HFr& a, const HFr& b) {
    HFr r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (fr_geq_mod(r.l)) fr_sub_mod(r.l);      // 2r < 2^256: no carry out
    return r;
}
CPG_HD HFr fr_sub(const HFr& a, const HFr& b) {
    HFr r; u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)br; r.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + FR_MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
CPG_HD HFr fr_mul(const HFr& a, const HFr& b) {
    uint64_t t[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; uint64_t t5 = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = ((u128)m * FR_MOD[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * FR_MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t5 + (uint64_t)(c >> 64);
    }
    HFr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fr_geq_mod(r.l)) fr_sub_mod(r.l);
    return r;
}
CPG_HD HFr fr_zero() { HFr r = {{0, 0, 0, 0}}; return r; }
CPG_HD HFr fr_one() { return fr_r1(); }
CPG_HD bool fr_is_zero(const HFr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
CPG_HD bool fr_eq(const HFr& a, const HFr& b) { return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3]; }
CPG_HD HFr fr_neg(const HFr& a) { return fr_sub(fr_zero(), a); }
CPG_HD HFr fr_from_u64(uint64_t v) { HFr r = {{v, 0, 0, 0}}; return fr_mul(r, fr_r2()); }
// canonical 32-byte little-endian -> HFr; false if >= r
CPG_HD bool fr_from_bytes(HFr* out, const uint8_t* b) {
    HFr r; copy32(r.l, b);
    if (fr_geq_mod(r.l)) return false;
    *out = fr_mul(r, fr_r2());
    return true;
}
CPG_HD void fr_to_bytes(uint8_t* b, const HFr& a) {
    HFr one = {{1, 0, 0, 0}};
    HFr r = fr_mul(a, one);
    copy32(b, r.l);
}
CPG_HD HFr fr_pow_u64(HFr a, uint64_t e) {
    HFr r = fr_one();
    while (e) { if (e & 1) r = fr_mul(r, a); a = fr_mul(a, a); e >>= 1; }
    return r;
}
// Inverse by the binary extended Euclidean algorithm (HAC 14.61) on the plain integers: ~600 iterations of
// shifts / subtractions on 4 limbs instead of the 420 Montgomery products of a^(r-2) - the per-round challenge
// inversions were half of the prover's per-proof Fr work.  0 -> 0.
//   input A = a R (Montgomery form);  binv(A) = a^-1 R^-1;  mont(binv(A), R^3) = a^-1 R.
CPG_HD bool fr_u256_geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
    return true;
}
CPG_HD void fr_u256_sub(uint64_t* a, const uint64_t* b) {            // a -= b (a >= b)
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - (uint64_t)br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
CPG_HD void fr_u256_shr1(uint64_t* a) {
    for (int i = 0; i < 3; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[3] >>= 1;
}
CPG_HD void fr_halve_mod(uint64_t* x) {                              // x / 2 mod r, x < r: (x or x + r) >> 1, x + r < 2^256
    if (x[0] & 1) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)x[i] + FR_MOD[i]; x[i] = (uint64_t)c; c >>= 64; } }
    fr_u256_shr1(x);
}
CPG_HD void fr_sub_into_mod(uint64_t* x, const uint64_t* y) {        // x = x - y mod r, both < r
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - y[i] - (uint64_t)br; x[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)x[i] + FR_MOD[i]; x[i] = (uint64_t)c; c >>= 64; } }
}
CPGH_CALL HFr fr_inv(HFr a) {
    if (fr_is_zero(a)) return a;
    uint64_t u[4], v[4], x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) { u[i] = a.l[i]; v[i] = FR_MOD[i]; }
    for (;;) {                                                       // gcd(u, v) = 1 (r prime): ends with u = 1 or v = 1
        const bool u1 = u[0] == 1 && (u[1] | u[2] | u[3]) == 0, v1 = v[0] == 1 && (v[1] | v[2] | v[3]) == 0;
        if (u1 || v1) {
            HFr t, r3;
            for (int i = 0; i < 4; i++) { t.l[i] = u1 ? x1[i] : x2[i]; r3.l[i] = CPGH_SEL(FR_R3)[i]; }
            return fr_mul(t, r3);
        }
        if (!(u[0] & 1)) { fr_u256_shr1(u); fr_halve_mod(x1); }
        else if (!(v[0] & 1)) { fr_u256_shr1(v); fr_halve_mod(x2); }
        else if (fr_u256_geq(u, v)) { fr_u256_sub(u, v); fr_sub_into_mod(x1, x2); }
        else { fr_u256_sub(v, u); fr_sub_into_mod(x2, x1); }
    }
}
// in-place batch inversion (Montgomery's trick); zeros stay zero
CPG_HD void fr_batch_inv(HFr* v, size_t n, HFr* scratch) {
    HFr acc = fr_one();
    for (size_t i = 0; i < n; i++) { scratch[i] = acc; if (!fr_is_zero(v[i])) acc = fr_mul(acc, v[i]); }
    acc = fr_inv(acc);
    for (size_t i = n; i-- > 0;) {
        if (fr_is_zero(v[i])) continue;
        HFr t = fr_mul(acc, scratch[i]);
        acc = fr_mul(acc, v[i]);
        v[i] = t;
    }
}

// the identity's 48-byte encoding: appended for the blinder slots of T_wb / U_wb (cp/curdleproofs.py:124-136)
#define CPGH_INF48_INIT {0xc0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
static const uint8_t H_INF48[48] = CPGH_INF48_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint8_t D_INF48[48] = CPGH_INF48_INIT;
#endif

// -------------------------------------------------------------- Merlin + challenges ---
// Labels are string literals; their length is taken at compile time (no strlen in device code).
struct Transcript {
    Strobe128 s;
    template <size_t N> CPG_HD void init(const char (&label)[N]) {
        const uint8_t merlin[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
        s.init(merlin, 11);
        const char ds[8] = "dom-sep";
        append(ds, (const uint8_t*)label, N - 1);
    }
    // MerlinTranscript.append_message / challenge_bytes with run-time labels (merlin_transcript.py:11-24)
    CPG_HD void append_rt(const uint8_t* label, size_t nl, const uint8_t* msg, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s.meta_ad(label, nl, false);
        s.meta_ad(len, 4, true);
        s.ad(msg, n, false);
    }
    CPG_HD void challenge_bytes_rt(const uint8_t* label, size_t nl, uint8_t* out, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s.meta_ad(label, nl, false);
        s.meta_ad(len, 4, true);
        s.prf(out, n, false);
    }
    CPG_HD void init_rt(const uint8_t* label, size_t nl) {
        const uint8_t merlin[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
        s.init(merlin, 11);
        const uint8_t ds[7] = {'d', 'o', 'm', '-', 's', 'e', 'p'};
        append_rt(ds, 7, label, nl);
    }
    template <size_t N> CPG_HD void append(const char (&label)[N], const uint8_t* msg, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s.meta_ad((const uint8_t*)label, N - 1, false);
        s.meta_ad(len, 4, true);
        s.ad(msg, n, false);
    }
    template <size_t N> CPG_HD void append_point(const char (&label)[N], const uint8_t* p48) { append(label, p48, 48); }
    template <size_t N> CPG_HD void append_fr(const char (&label)[N], const HFr& v) { uint8_t b[32]; fr_to_bytes(b, v); append(label, b, 32); }
    template <size_t N> CPG_HD void challenge_bytes(const char (&label)[N], uint8_t* out, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        s.meta_ad((const uint8_t*)label, N - 1, false);
        s.meta_ad(len, 4, true);
        s.prf(out, n, false);
    }
    // rejection-sample a non-zero scalar < r, then bind it back (curdleproofs_transcript.py:15-25)
    template <size_t N> CPG_HD HFr challenge(const char (&label)[N]) {
        for (;;) {
            uint8_t raw[32];
            challenge_bytes(label, raw, 32);
            HFr v;
            if (!fr_from_bytes(&v, raw) || fr_is_zero(v)) continue;
            append(label, raw, 32);
            return v;
        }
    }
};

// A transcript driven by a byte script - how the KAT tests (and any host language) reach this implementation through
// the C ABI (cpg_merlin_script).  Record: op u8 | more u8 | label_len u16 | n u32 | label | data[n] (no data for the
// two output ops, where n = bytes wanted).  Returns the number of output bytes, or (size_t)-1 on a malformed script.
enum { MS_STROBE_INIT = 0, MS_META_AD = 1, MS_AD = 2, MS_PRF = 3, MS_KEY = 4, MS_MERLIN_INIT = 5, MS_MERLIN_APPEND = 6, MS_MERLIN_CHALLENGE = 7 };
CPG_HD size_t merlin_run_script(const uint8_t* sc, size_t len, uint8_t* out, size_t cap, uint32_t warp = 0) {
    Transcript tr;
    memset(&tr, 0, sizeof tr);
    tr.s.warp = (uint8_t)warp;                 // device: the 32 lanes of a warp run this very call in lock-step
    size_t p = 0, o = 0;
    while (p < len) {
        if (p + 8 > len) return (size_t)-1;
        const uint8_t op = sc[p], more = sc[p + 1];
        const size_t nl = (size_t)sc[p + 2] | ((size_t)sc[p + 3] << 8);
        const size_t n = (size_t)sc[p + 4] | ((size_t)sc[p + 5] << 8) | ((size_t)sc[p + 6] << 16) | ((size_t)sc[p + 7] << 24);
        p += 8;
        const bool emits = op == MS_PRF || op == MS_MERLIN_CHALLENGE;
        if (p + nl + (emits ? 0 : n) > len || (emits && o + n > cap)) return (size_t)-1;
        const uint8_t* label = sc + p; p += nl;
        const uint8_t* data = sc + p; if (!emits) p += n;
        switch (op) {
        case MS_STROBE_INIT: tr.s.init(data, n); break;
        case MS_META_AD: tr.s.meta_ad(data, n, more != 0); break;
        case MS_AD: tr.s.ad(data, n, more != 0); break;
        case MS_KEY: tr.s.key(data, n, more != 0); break;
        case MS_PRF: tr.s.prf(out + o, n, more != 0); o += n; break;
        case MS_MERLIN_INIT: tr.init_rt(data, n); break;
        case MS_MERLIN_APPEND: tr.append_rt(label, nl, data, n); break;
        case MS_MERLIN_CHALLENGE: tr.challenge_bytes_rt(label, nl, out + o, n); o += n; break;
        default: return (size_t)-1;
        }
    }
    return o;
}

}  // namespace cpgh

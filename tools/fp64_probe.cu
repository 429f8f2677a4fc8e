// FP64-pipe Fq product (csrc/fq52.cuh) against the integer-pipe one (csrc/field.cuh): correctness on random
// operands and edge values, then throughput alone, and with both kinds of warps resident on the same SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../curdleproofs_pie_b200/csrc/field.cuh"
#include "../curdleproofs_pie_b200/csrc/fq52.cuh"
using namespace cpg;

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

// 12 x u32 plain integer -> 8 x 52-bit limbs
__host__ __device__ inline void repack_32_to_52(uint64_t (&o)[8], const uint32_t* l) {
    for (int j = 0; j < 8; j++) {
        uint64_t v = 0;
        for (int b = 0; b < 52; b++) {
            int bit = 52 * j + b;
            if (bit < 384 && ((l[bit >> 5] >> (bit & 31)) & 1u)) v |= 1ull << b;
        }
        o[j] = v;
    }
}
__host__ __device__ inline void repack_52_to_32(uint32_t* l, const uint64_t (&x)[8]) {
    for (int i = 0; i < 12; i++) l[i] = 0;
    for (int bit = 0; bit < 384; bit++)
        if ((x[bit / 52] >> (bit % 52)) & 1ull) l[bit >> 5] |= 1u << (bit & 31);
}

// out_int = x*y mod p through the integer pipe, out_fp = the same through the FP64 pipe (both plain, canonical);
// also x^2 through ms52 and a 20-long mixed chain
__global__ void k_check(const uint32_t* xs, const uint32_t* ys, uint32_t* out_int, uint32_t* out_fp, uint32_t* out_int_chain, uint32_t* out_fp_chain, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Fq x, y;
    for (int i = 0; i < 12; i++) { x.l[i] = xs[t * 12 + i]; y.l[i] = ys[t * 12 + i]; }
    Fq xm = to_mont(x), ym = to_mont(y);
    Fq zi = from_mont(mul(xm, ym));
    for (int i = 0; i < 12; i++) out_int[t * 12 + i] = zi.l[i];
    Fq ci = xm;
    for (int i = 0; i < 10; i++) { ci = sqr(ci); ci = mul(ci, ym); }
    ci = from_mont(ci);
    for (int i = 0; i < 12; i++) out_int_chain[t * 12 + i] = ci.l[i];
    // FP64 path: the integer Montgomery form (x 2^384) enters by one product with 2^448
    uint64_t l52[8];
    F52 a, b, k1, k2;
    repack_32_to_52(l52, xm.l);
    for (int j = 0; j < 8; j++) a.v[j] = (double)l52[j];
    repack_32_to_52(l52, ym.l);
    for (int j = 0; j < 8; j++) b.v[j] = (double)l52[j];
    for (int j = 0; j < 8; j++) { k1.v[j] = (double)D_C384_416[j]; k2.v[j] = (double)D_C416_384[j]; }
    F52 af = mul(a, k1), bf = mul(b, k1);
    F52 zf = mul(mul(af, bf), k2);                       // x y 2^384: the integer Montgomery form again
    uint64_t o[8];
    fq52_canon(o, zf);
    Fq back;
    repack_52_to_32(back.l, o);
    back = from_mont(back);
    for (int i = 0; i < 12; i++) out_fp[t * 12 + i] = back.l[i];
    F52 cf = af;
    for (int i = 0; i < 10; i++) { cf = sqr(cf); cf = mul(cf, bf); }
    cf = mul(cf, k2);
    fq52_canon(o, cf);
    repack_52_to_32(back.l, o);
    back = from_mont(back);
    for (int i = 0; i < 12; i++) out_fp_chain[t * 12 + i] = back.l[i];
}

// mode 0: every warp runs the integer chain; 1: every warp the FP64 chain; 2: even warps integer, odd warps FP64;
// 3 / 4 / 5: the same with squarings
__global__ void __launch_bounds__(128, 3) k_chain(int mode, int iters, const uint32_t* seed, uint32_t* sink) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int warp = threadIdx.x >> 5;
    bool sq = mode >= 3;
    int m = sq ? mode - 3 : mode;
    bool fp = m == 1 || (m == 2 && (warp & 1));
    Fq x;
    for (int i = 0; i < 12; i++) x.l[i] = seed[i] + (i == 0 ? (uint32_t)t : 0u);
    x.l[11] &= 0x0fffffffu;
    if (!fp) {
        Fq a = x, b = x;
        b.l[0] ^= 5;
        if (sq) for (int i = 0; i < iters; i++) a = sqr(a);
        else for (int i = 0; i < iters; i++) a = mul(a, b);
        uint32_t s = 0;
        for (int i = 0; i < 12; i++) s ^= a.l[i];
        if (s == 0x12345678u) sink[0] = s;
    } else {
        uint64_t l52[8];
        repack_32_to_52(l52, x.l);
        F52 a, b;
        for (int j = 0; j < 8; j++) { a.v[j] = (double)l52[j]; b.v[j] = (double)(l52[j] ^ 5); }
        if (sq) {
#pragma unroll 1
            for (int i = 0; i < iters; i++) a = sqr(a);
        } else {
#pragma unroll 1
            for (int i = 0; i < iters; i++) a = mul(a, b);
        }
        double s = 0;
        for (int j = 0; j < 8; j++) s += a.v[j];
        if (s == 1.25) sink[1] = 1;
    }
}

// the FP64 chain alone at MINB resident blocks of 128 threads (the integer product needs 168 registers = 3 blocks;
// this one fits 80 registers = 6 blocks without spilling)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_fp(int sq, int iters, const uint32_t* seed, uint32_t* sink) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t xl[12];
    for (int i = 0; i < 12; i++) xl[i] = seed[i] + (i == 0 ? (uint32_t)t : 0u);
    xl[11] &= 0x0fffffffu;
    uint64_t l52[8];
    repack_32_to_52(l52, xl);
    F52 a, b;
    for (int j = 0; j < 8; j++) { a.v[j] = (double)l52[j]; b.v[j] = (double)(l52[j] ^ 5); }
    if (sq) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) a = sqr(a);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) a = mul(a, b);
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += a.v[j];
    if (s == 1.25) sink[1] = 1;
}
template <int MINB>
static int run_fp(int sms, const uint32_t* dx, uint32_t* sink) {
    for (int sq = 0; sq < 2; sq++) {
        int iters = 2000, blocks = sms * MINB;
        k_fp<MINB><<<blocks, 128>>>(sq, 10, dx, sink);
        CKC(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_fp<MINB><<<blocks, 128>>>(sq, iters, dx, sink);
        cudaEventRecord(e1);
        CKC(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s FP64 pipe, %d blocks of 128 per SM   %8.3f ms  %.4g products/s\n", sq ? "sqr:" : "mul:", MINB, ms, (double)blocks * 128 * iters / (ms * 1e-3));
    }
    return 0;
}

static void rnd_below_p(uint32_t* l) {
    for (;;) {
        for (int i = 0; i < 12; i++) l[i] = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
        l[11] &= 0x1fffffffu;
        bool ge = true;
        for (int i = 11; i >= 0; i--) { if (l[i] != H_FQ_P[i]) { ge = l[i] > H_FQ_P[i]; break; } }
        if (!ge) return;
    }
}

int main() {
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, clk);
    const int n = 4096;
    uint32_t *hx = (uint32_t*)malloc(n * 48), *hy = (uint32_t*)malloc(n * 48);
    srand(12345);
    for (int t = 0; t < n; t++) { rnd_below_p(hx + t * 12); rnd_below_p(hy + t * 12); }
    // edge operands: 0, 1, p-1, all-ones limbs below p
    for (int i = 0; i < 12; i++) { hx[i] = 0; hx[12 + i] = i == 0; hx[24 + i] = H_FQ_P[i] - (i == 0); hy[24 + i] = H_FQ_P[i] - (i == 0); hx[36 + i] = i == 11 ? 0x0fffffffu : 0xffffffffu; hy[36 + i] = hx[36 + i]; }
    uint32_t *dx, *dy, *o1, *o2, *o3, *o4;
    CKC(cudaMalloc(&dx, n * 48)); CKC(cudaMalloc(&dy, n * 48)); CKC(cudaMalloc(&o1, n * 48)); CKC(cudaMalloc(&o2, n * 48)); CKC(cudaMalloc(&o3, n * 48)); CKC(cudaMalloc(&o4, n * 48));
    CKC(cudaMemcpy(dx, hx, n * 48, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dy, hy, n * 48, cudaMemcpyHostToDevice));
    k_check<<<n / 128, 128>>>(dx, dy, o1, o2, o3, o4, n);
    CKC(cudaDeviceSynchronize());
    uint32_t *h1 = (uint32_t*)malloc(n * 48), *h2 = (uint32_t*)malloc(n * 48), *h3 = (uint32_t*)malloc(n * 48), *h4 = (uint32_t*)malloc(n * 48);
    CKC(cudaMemcpy(h1, o1, n * 48, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(h2, o2, n * 48, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(h3, o3, n * 48, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(h4, o4, n * 48, cudaMemcpyDeviceToHost));
    int bad = 0, badc = 0;
    for (int t = 0; t < n; t++) {
        for (int i = 0; i < 12; i++) { if (h1[t * 12 + i] != h2[t * 12 + i]) { bad++; break; } }
        for (int i = 0; i < 12; i++) { if (h3[t * 12 + i] != h4[t * 12 + i]) { badc++; break; } }
    }
    printf("correctness: %d operand pairs, product mismatches %d, 20-step sqr/mul chain mismatches %d\n", n, bad, badc);
    if (bad || badc) return 2;
    uint32_t* sink;
    CKC(cudaMalloc(&sink, 64));
    const char* names[6] = {"mul: integer pipe only", "mul: FP64 pipe only", "mul: half the warps each", "sqr: integer pipe only", "sqr: FP64 pipe only", "sqr: half the warps each"};
    for (int blocks_per_sm = 3; blocks_per_sm <= 4; blocks_per_sm++)
    for (int mode = 0; mode < 6; mode++) {
        int iters = 2000, blocks = prop.multiProcessorCount * blocks_per_sm;
        k_chain<<<blocks, 128>>>(mode, 10, dx, sink);
        CKC(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_chain<<<blocks, 128>>>(mode, iters, dx, sink);
        cudaEventRecord(e1);
        CKC(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * 128 * iters;
        printf("%-28s blocks/SM=%d  %8.3f ms  %.4g products/s\n", names[mode], blocks_per_sm, ms, ops / (ms * 1e-3));
    }
    run_fp<3>(prop.multiProcessorCount, dx, sink);
    run_fp<4>(prop.multiProcessorCount, dx, sink);
    run_fp<5>(prop.multiProcessorCount, dx, sink);
    run_fp<6>(prop.multiProcessorCount, dx, sink);
    run_fp<8>(prop.multiProcessorCount, dx, sink);
    return 0;
}

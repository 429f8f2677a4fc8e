// Integer-pipe microbenchmarks for sm_100a: what bounds a 12x32-bit Montgomery product?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/imad_probe tools/imad_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../curdleproofs_pie_b200/csrc/field.cuh"
using namespace cpg;

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

// IMAD.WIDE.U32 with a data-dependent multiplicand (acc.lo), so nothing is loop-invariant:
// (lo,hi) = a * lo + (lo,hi).  8 independent chains per thread.
#define WIDE_DEP(lo, hi, a) asm volatile("{ .reg .u32 t; mov.u32 t, %0; mad.lo.cc.u32 %0, %2, t, %0; madc.hi.u32 %1, %2, t, %1; }" : "+r"(lo), "+r"(hi) : "r"(a))
__global__ void __launch_bounds__(256) k_wide_same(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a = a0 + threadIdx.x;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = j + threadIdx.x + b0; hi[j] = j; }
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) WIDE_DEP(lo[j], hi[j], a);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
    if (s == 0x12345678u) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_wide_distinct(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a[8];
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = j + threadIdx.x + b0; hi[j] = j; a[j] = a0 * (j + 1) + threadIdx.x; }
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) WIDE_DEP(lo[j], hi[j], a[j]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
    if (s == 0x12345678u) sink[0] = s;
}
// 32-bit IMAD (lo) and IMAD.HI with data-dependent operands
__global__ void __launch_bounds__(256) k_imad32(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a = a0 + threadIdx.x;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = j + threadIdx.x + b0;
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %1, %0, %0;" : "+r"(x[j]) : "r"(a));
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    if (s == 0x12345678u) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_imadhi(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a = a0 + threadIdx.x;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = j + threadIdx.x + b0;
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mad.hi.u32 %0, %1, %0, %0;" : "+r"(x[j]) : "r"(a));
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    if (s == 0x12345678u) sink[0] = s;
}
// carry-chained rows: exactly cmad_even<12> (6 IMAD.WIDE.U32.X per row), CH independent accumulators
template <int CH>
__global__ void __launch_bounds__(256) k_widex_rows(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a[12], b = b0 + blockIdx.x;
    uint32_t acc[CH][12];
#pragma unroll
    for (int j = 0; j < 12; j++) { a[j] = a0 * (j + 1) + threadIdx.x; }
#pragma unroll
    for (int c = 0; c < CH; c++)
#pragma unroll
        for (int j = 0; j < 12; j++) acc[c][j] = j + c;
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) cmad_even<12>(acc[c], a, acc[c][11] + b);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++)
#pragma unroll
        for (int j = 0; j < 12; j++) s ^= acc[c][j];
    if (s == 0x12345678u) sink[0] = s;
}
// the library's Montgomery product, CH independent dependent-chains per thread
template <int CH, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_mont(uint64_t iters, uint64_t* sink) {
    Fq x[CH], y[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { x[c] = Fq::one(); y[c] = Fq::one(); x[c].l[0] += threadIdx.x + c; y[c].l[1] += blockIdx.x + c; }
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = mul(x[c], y[c]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s ^= x[c].l[0] ^ x[c].l[5];
    if (s == 0x12345678u) sink[0] = s;
}
// ALU pipe alone (IADD3 carry chains) and mixed with IMAD.WIDE, to see whether the pipes overlap
__global__ void __launch_bounds__(256) k_iadd_chain(uint64_t iters, uint32_t a0, uint64_t* sink) {
    uint32_t x[12], y[12];
#pragma unroll
    for (int j = 0; j < 12; j++) { x[j] = a0 + j + threadIdx.x; y[j] = a0 * j + blockIdx.x; }
    for (uint64_t i = 0; i < iters; i++) {
        x[0] = add_cc(x[0], y[0]);
#pragma unroll
        for (int j = 1; j < 11; j++) x[j] = addc_cc(x[j], y[j]);
        x[11] = addc(x[11], y[11]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) s ^= x[j];
    if (s == 0x12345678u) sink[0] = s;
}

template <class K, class... A>
double run(const char* name, double ops_per_thread_iter, uint64_t iters, int blocks, int threads, K kern, A... args) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        kern<<<blocks, threads>>>(iters, args...);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t e = cudaGetLastError();
    double rate = ops_per_thread_iter * iters * blocks * (double)threads / (ms * 1e-3);
    printf("%-44s blocks=%5d thr=%3d  %8.3f ms  %.4g ops/s %s\n", name, blocks, threads, ms, rate, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return rate;
}

int main() {
    cudaDeviceProp p; CKC(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    uint64_t* sink; CKC(cudaMalloc(&sink, 8));
    double peak = run("mad.wide same operands (MAC)", 8, 20000, sms * 8, 256, k_wide_same, 12345u, 6789u, sink);
    run("mad.wide distinct operands (MAC)", 8, 20000, sms * 8, 256, k_wide_distinct, 12345u, 6789u, sink);
    run("IMAD (32-bit lo) dependent (MAC)", 8, 20000, sms * 8, 256, k_imad32, 12345u, 6789u, sink);
    run("IMAD.HI dependent (MAC)", 8, 20000, sms * 8, 256, k_imadhi, 12345u, 6789u, sink);
    run("IMAD.WIDE.X rows, 1 chain (MAC)", 6, 20000, sms * 8, 256, k_widex_rows<1>, 12345u, 6789u, sink);
    run("IMAD.WIDE.X rows, 2 chains (MAC)", 12, 20000, sms * 8, 256, k_widex_rows<2>, 12345u, 6789u, sink);
    run("IMAD.WIDE.X rows, 4 chains (MAC)", 24, 10000, sms * 8, 256, k_widex_rows<4>, 12345u, 6789u, sink);
    run("IADD3.X 12-limb add chains (adds)", 12, 20000, sms * 8, 256, k_iadd_chain, 12345u, sink);
    double m1 = run("mont_mul 1 chain/thread, 256thr (modmul)", 1, 2000, sms * 8, 256, k_mont<1, 256>, sink);
    run("mont_mul 2 chains/thread, 256thr (modmul)", 2, 2000, sms * 8, 256, k_mont<2, 256>, sink);
    run("mont_mul 4 chains/thread, 128thr (modmul)", 4, 1000, sms * 8, 128, k_mont<4, 128>, sink);
    run("mont_mul 1 chain, 4 warps/SM (modmul)", 1, 2000, sms, 128, k_mont<1, 128>, sink);
    run("mont_mul 1 chain, 8 warps/SM (modmul)", 1, 2000, sms * 2, 128, k_mont<1, 128>, sink);
    run("mont_mul 1 chain, 16 warps/SM (modmul)", 1, 2000, sms * 4, 128, k_mont<1, 128>, sink);
    run("mont_mul 1 chain, 32 warps/SM (modmul)", 1, 2000, sms * 8, 128, k_mont<1, 128>, sink);
    printf("ideal modmul/s at 300 MAC = %.4g ; mont_mul achieves %.1f%%\n", peak / 300, 100 * m1 / (peak / 300));
    return 0;
}

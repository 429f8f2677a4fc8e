// Do the FP64 pipe (DFMA), the multiply pipe (IMAD.WIDE.U32) and the ALU pipe (IADD3) of an sm_100a SM issue
// side by side?  Eight independent chains of each kind per thread, alone and interleaved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define WIDE_DEP(lo, hi, a) asm volatile("{ .reg .u32 t; mov.u32 t, %0; mad.lo.cc.u32 %0, %2, t, %0; madc.hi.u32 %1, %2, t, %1; }" : "+r"(lo), "+r"(hi) : "r"(a))
#define DFMA_DEP(x, a, b) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b))
#define DADD_DEP(x, a) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(x) : "d"(a))
#define IADD_DEP(x, y, a) asm volatile("{ add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2; }" : "+r"(x), "+r"(y) : "r"(a))

template <int MODE>   // bit 0: IMAD.WIDE, bit 1: DFMA, bit 2: IADD3 pairs, bit 3: DADD, bit 4: 64-bit three-input adds (IADD3 with two carries + IADD3.X)
__global__ void __launch_bounds__(256) k(uint64_t iters, uint32_t a0, double d0, uint64_t* sink) {
    uint32_t a = a0 + threadIdx.x;
    uint32_t lo[8], hi[8], x[8], y[8];
    uint64_t w[8], wa = a0 * 77ull + threadIdx.x, wb = a0 + 5ull;
    double g[8];
    double f[8], fa = d0 + 1e-9 * threadIdx.x, fb = d0 * 0.5;
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = j + threadIdx.x; hi[j] = j; x[j] = j * 3 + threadIdx.x; y[j] = j; f[j] = 1.0 + j; g[j] = 2.0 + j; w[j] = j; }
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE & 1) WIDE_DEP(lo[j], hi[j], a);
            if (MODE & 2) DFMA_DEP(f[j], fa, fb);
            if (MODE & 4) IADD_DEP(x[j], y[j], a);
            if (MODE & 8) DADD_DEP(g[j], fa);
            if (MODE & 16) { w[j] = w[j] + wa + wb; asm volatile("" : "+l"(w[j])); }
        }
    }
    uint32_t s = 0;
    double t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { s ^= lo[j] ^ hi[j] ^ x[j] ^ y[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32); t += f[j] + g[j]; }
    if (s == 0x12345678u || t == 1.2345) sink[0] = s;
}

template <int MODE>
static void run(const char* name, int sms) {
    uint64_t* sink;
    cudaMalloc(&sink, 8);
    uint64_t iters = 20000;
    int blocks = sms * 4;
    k<MODE><<<blocks, 256>>>(10, 3, 0.999, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(iters, 3, 0.999, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per_kind = (double)blocks * 256 * iters * 8;
    printf("%-44s %8.3f ms   %.1f lanes/clk/SM of EACH kind\n", name, ms, per_kind / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    int s = prop.multiProcessorCount;
    run<1>("IMAD.WIDE.U32 alone", s);
    run<2>("DFMA alone", s);
    run<4>("IADD3 + IADD3.X pair alone (pairs)", s);
    run<3>("IMAD.WIDE + DFMA interleaved", s);
    run<5>("IMAD.WIDE + IADD3 pair interleaved", s);
    run<6>("DFMA + IADD3 pair interleaved", s);
    run<7>("IMAD.WIDE + DFMA + IADD3 pair interleaved", s);
    run<8>("DADD alone", s);
    run<10>("DFMA + DADD interleaved", s);
    run<16>("64-bit 3-input add alone (IADD3 P0,P1 + IADD3.X)", s);
    run<18>("DFMA + 64-bit 3-input add", s);
    run<26>("DFMA + DADD + 64-bit 3-input add", s);
    return 0;
}

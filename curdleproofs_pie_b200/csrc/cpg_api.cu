// libcpg.so - C ABI (include/cpg.h) over the sm_100a kernels.  One functor (msm.cuh) = one kernel.
//
// Built two ways from this one file:
//   nvcc -gencode arch=compute_100a,code=sm_100a  -> curdleproofs_pie_b200/lib/libcpg.so   (the product)
//   g++ -x c++ -DCPG_HOST_EMU                     -> tests/_build/libcpg_hostseam.so       (TEST SEAM ONLY:
//        every "kernel" becomes a host loop over the same per-thread functor with the PTX carry flag
//        emulated, so the CPU-only test tier can exercise the exact launch logic; the product package
//        never loads it and refuses to run without a CUDA device)
#include "../../include/cpg.h"
#include "msm.cuh"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#ifndef CPG_HOST_EMU
#include <cuda_runtime.h>
#endif

using namespace cpg;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
int g_device = -1;
int g_msm_path = 0;          // cpg_msm_force_path: 0 = by shape, 1 = per-(msm, window) threads, 2 = per-term threads
int g_msm_affine = 0;        // cpg_msm_set_accumulate: 0 = mixed XYZZ additions in the bucket accumulation (default: measured faster), 1 = batched affine additions
std::mutex g_init_mu;
Jac* g_generator = nullptr;  // device copy of the generator (Jacobian)

int fail(const std::string& msg) { g_err = msg; return 1; }

#ifndef CPG_HOST_EMU
cudaStream_t g_stream = nullptr;       // library default stream
thread_local cudaStream_t t_stream = nullptr;
thread_local bool t_stream_set = false;
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
cudaStream_t cur() { return t_stream_set ? t_stream : g_stream; }
// hooks of the pipelined single MSM (msm_large_pipelined): the per-term MSM pipeline waits on t_pipe_wait before its first
// kernel and records t_pipe_signal right after its bucket accumulation
thread_local cudaEvent_t t_pipe_signal = nullptr;
thread_local cudaStream_t t_pipe_tail = nullptr;      // high-priority stream the window reduction moves to (pipelined MSM)

int ck(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(x) do { if (int rc_ = ck((x), #x)) return rc_; } while (0)

template <class F, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_each(const F f, uint64_t n) {
    uint64_t t = blockIdx.x * (uint64_t)BLOCK + threadIdx.x;
    if (t < n) f(t);
}
// SortDigits with its per-thread histogram (NB+1 words) and rank (NB halfwords) arrays in shared
// memory: element k of thread t sits at [k*BLOCK + t], so a thread always stays in its own bank.
constexpr int SORT_BLOCK = 128;
__global__ void __launch_bounds__(SORT_BLOCK) k_sort_digits_smem(SortDigits f, uint64_t n_threads) {
    extern __shared__ uint32_t sort_smem[];
    const MsmShape& s = f.s;
    uint32_t* off_base = sort_smem + threadIdx.x;
    uint16_t* rk_base = (uint16_t*)(sort_smem + (size_t)(s.NB + 1) * SORT_BLOCK) + threadIdx.x;
    uint64_t t = blockIdx.x * (uint64_t)SORT_BLOCK + threadIdx.x;
    if (t >= n_threads) return;
    uint32_t m = (uint32_t)(t / s.wn), w = s.w0 + (uint32_t)(t % s.wn);
    StridedView<uint32_t> off{off_base, SORT_BLOCK};
    StridedView<uint16_t> rk{rk_base, SORT_BLOCK};
    sort_digits_body(s, f.dig + (uint64_t)m * s.n * s.W + w, f.sorted + t * (uint64_t)s.n, off, rk, f.rank != nullptr);
    uint32_t* goff = f.boff + t * (uint64_t)(s.NB + 1);
    for (uint32_t b = 0; b <= s.NB; b++) goff[b] = off[b];
    if (f.rank) {
        uint16_t* grk = f.rank + t * (uint64_t)s.NB;
        for (uint32_t b = 0; b < s.NB; b++) grk[b] = rk[b];
    }
}

// BucketAccumulateAffine: every thread of the grid runs (the padding threads of the last block too, owning no buckets),
// because its lanes vote on the loop exits (msm.cuh)
__global__ void __launch_bounds__(128, 3) k_bucket_affine(const BucketAccumulateAffine f, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)128 + threadIdx.x;
    f.run(t, t < n);
}

// ---- HornerJac for FEW MSMs: the 255 dependent doublings of one MSM's Horner pass on one thread are
// 1.8 ms of pure latency (7 dependent Fq products per doubling, 0.9 us each).  Here a warp owns an MSM and
// four of its lanes each take ONE of the independent products of a stage; the results travel by shuffle:
//   doubling (dbl-2009-l, a = 0):  {X^2, Y^2, Y Z} -> {B^2, (X+B)^2, (3A)^2} -> {E (D - X3)}      3 stages, not 7
//   addition (add-2007-bl):        4 -> 4 -> 2 -> 4 -> 2 products                                   5 stages, not 16
// Every lane keeps the whole accumulator, so control flow (identity / equal-point cases) stays warp-uniform.
__device__ __forceinline__ Fq warp_bcast(const Fq& v, int src) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
__device__ __forceinline__ Fq pick4(int lane, const Fq& a, const Fq& b, const Fq& c, const Fq& d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = lane == 0 ? a.l[i] : lane == 1 ? b.l[i] : lane == 2 ? c.l[i] : d.l[i];
    return r;
}
__device__ Jac coop_dbl(const Jac& p, int lane) {
    if (is_inf(p)) return p;
    Fq r1 = mul(pick4(lane, p.X, p.Y, p.Y, p.X), pick4(lane, p.X, p.Y, p.Z, p.X));        // A = X^2 | B = Y^2 | Y Z
    Fq A = warp_bcast(r1, 0), B = warp_bcast(r1, 1), YZ = warp_bcast(r1, 2);
    Fq E = add(dbl(A), A), t = add(p.X, B);
    Fq o2 = pick4(lane, B, t, E, B);
    Fq r2 = mul(o2, o2);                                                                    // C = B^2 | t^2 | F = E^2
    Fq C = warp_bcast(r2, 0), T2 = warp_bcast(r2, 1), F = warp_bcast(r2, 2);
    Fq D = dbl(sub(sub(T2, A), C));
    Jac r;
    r.X = sub(F, dbl(D));
    r.Z = dbl(YZ);
    Fq m = warp_bcast(mul(E, sub(D, r.X)), 0);
    r.Y = sub(m, dbl(dbl(dbl(C))));
    return r;
}
__device__ Jac coop_add(const Jac& p, const Jac& q, int lane) {
    if (is_inf(p)) return q;
    if (is_inf(q)) return p;
    Fq r1 = mul(pick4(lane, p.Z, q.Z, p.Y, q.Y), pick4(lane, p.Z, q.Z, q.Z, p.Z));        // Z1Z1 | Z2Z2 | Y1 Z2 | Y2 Z1
    Fq Z1Z1 = warp_bcast(r1, 0), Z2Z2 = warp_bcast(r1, 1), Y1Z2 = warp_bcast(r1, 2), Y2Z1 = warp_bcast(r1, 3);
    Fq r2 = mul(pick4(lane, p.X, q.X, Y1Z2, Y2Z1), pick4(lane, Z2Z2, Z1Z1, Z2Z2, Z1Z1));  // U1 | U2 | S1 | S2
    Fq U1 = warp_bcast(r2, 0), U2 = warp_bcast(r2, 1), S1 = warp_bcast(r2, 2), S2 = warp_bcast(r2, 3);
    Fq H = sub(U2, U1), rr = sub(S2, S1);
    if (H.is_zero()) return rr.is_zero() ? coop_dbl(p, lane) : jac_inf();
    rr = dbl(rr);
    Fq o3 = pick4(lane, dbl(H), add(p.Z, q.Z), H, H);
    Fq r3 = mul(o3, o3);                                                                    // I = (2H)^2 | (Z1+Z2)^2
    Fq I = warp_bcast(r3, 0), ZS = warp_bcast(r3, 1);
    Fq r4 = mul(pick4(lane, H, U1, rr, sub(sub(ZS, Z1Z1), Z2Z2)), pick4(lane, I, I, rr, H)); // J | V | rr^2 | Z3
    Fq J = warp_bcast(r4, 0), V = warp_bcast(r4, 1), RR2 = warp_bcast(r4, 2);
    Jac r;
    r.Z = warp_bcast(r4, 3);
    r.X = sub(sub(RR2, J), dbl(V));
    Fq r5 = mul(pick4(lane, rr, S1, rr, rr), pick4(lane, sub(V, r.X), J, rr, rr));          // rr (V - X3) | S1 J
    r.Y = sub(warp_bcast(r5, 0), dbl(warp_bcast(r5, 1)));
    return r;
}
__global__ void __launch_bounds__(32) k_horner_jac_coop(HornerJac f, uint64_t n_msm) {
    const uint64_t m = blockIdx.x;
    if (m >= n_msm) return;
    const int lane = threadIdx.x;
    const Jac* ws = f.wsum + m * (uint64_t)f.W;
    uint32_t w = f.hi();
    Jac acc;
    if (f.cont) acc = f.out[m];
    else acc = ws[--w];
    while (w-- > f.w_lo) {
        for (uint32_t j = 0; j < f.c; j++) acc = coop_dbl(acc, lane);
        acc = coop_add(acc, ws[w], lane);
    }
    if (lane == 0) f.out[m] = acc;
}

// ---- optional per-kernel timing (cpg_profile_*): one CUDA event pair per launch on the launching
// stream, resolved lazily.  Off by default; bench.py turns it on to time the dominant kernel live.
struct ProfRec { const char* name; cudaEvent_t a, b; uint64_t threads; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;
// work counters of the table-lookup MSMs, gathered only while profiling: (term, window) pairs with a non-zero
// coefficient = mixed additions FixedMsmWindow / VarTableMsmWindow really perform (they skip zero coefficients)
unsigned long long* g_d_counters = nullptr;   // [0] fixed-base, [1] per-base tracker tables

template <int BLOCK = 128, int MINB = 1, class F>
int launch(const F& f, uint64_t n) {
    if (!n) return 0;
    uint64_t grid = (n + BLOCK - 1) / BLOCK;
    if (grid > 0x7fffffffULL) return fail("launch: grid too large");
    ProfRec rec{F::kName, nullptr, nullptr, n};
    if (g_prof_on) {
        cudaEventCreate(&rec.a); cudaEventCreate(&rec.b);
        cudaEventRecord(rec.a, cur());
    }
    k_each<F, BLOCK, MINB><<<(unsigned)grid, BLOCK, 0, cur()>>>(f, n);
    if (g_prof_on) {
        cudaEventRecord(rec.b, cur());
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(rec);
    }
    g_launches++;
    return ck(cudaGetLastError(), "kernel launch");
}
// Point-arithmetic kernels: 3 resident blocks per SM (<= 168 registers, 12 warps) hide the dependent
// IMAD.WIDE chains better than 2 blocks of 240 registers (Decompress 70.8 -> 65.3 ms per 4.8 M points,
// verification +8 %).  Not for the scalar-multiplication kernels (jac_mul spills at 168: ProveShuffle
// 122 -> 140 ms).  CPG_OCC1=1 restores the compiler's own register budget for A/B runs.
bool occ1() { static const bool v = getenv("CPG_OCC1") != nullptr; return v; }
template <class F>
int launch_occ(const F& f, uint64_t n) { return occ1() ? launch<128, 1>(f, n) : launch<128, 3>(f, n); }
// The unchecked Decompress is a chain of 376 squarings + ~100 products on ONE value (the 16-entry window table lives in
// local memory): 96 registers, 5 resident blocks per SM instead of the 3 of a point addition.  Measured on 4.8 M points
// (profiles/r02_ab_decomp_occ.txt): 71.2 ms at 168 registers / 3 blocks (when the kernel still carried the subgroup
// check's code), 68.4 at 108 / 4, 67.4 at 96 / 5, 67.5 at 80 / 6, 67.7 at 64 / 8.  CPG_DECOMP_MINB=3 keeps the old launch.
int decomp_minb() { static const int v = getenv("CPG_DECOMP_MINB") ? atoi(getenv("CPG_DECOMP_MINB")) : 5; return v; }
template <class F>
int launch_decomp(const F& f, uint64_t n) { return decomp_minb() == 5 ? launch<128, 5>(f, n) : launch_occ(f, n); }
int launch_sort_digits(const SortDigits& f, uint64_t n) {
    const MsmShape& s = f.s;
    size_t smem = (size_t)(s.NB + 1) * SORT_BLOCK * 4 + (size_t)s.NB * SORT_BLOCK * 2;
    if (smem > 200 * 1024) return launch(f, n);                 // very wide windows: per-thread global arrays
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (int rc = ck(cudaFuncSetAttribute(k_sort_digits_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return rc;
        configured = smem;
    }
    if (!n) return 0;
    uint64_t grid = (n + SORT_BLOCK - 1) / SORT_BLOCK;
    ProfRec rec{SortDigits::kName, nullptr, nullptr, n};
    if (g_prof_on) { cudaEventCreate(&rec.a); cudaEventCreate(&rec.b); cudaEventRecord(rec.a, cur()); }
    k_sort_digits_smem<<<(unsigned)grid, SORT_BLOCK, smem, cur()>>>(f, n);
    if (g_prof_on) { cudaEventRecord(rec.b, cur()); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back(rec); }
    g_launches++;
    return ck(cudaGetLastError(), "kernel launch");
}
int launch_bucket_affine(const BucketAccumulateAffine& f, uint64_t n) {
    if (!n) return 0;
    ProfRec rec{BucketAccumulateAffine::kName, nullptr, nullptr, n};
    if (g_prof_on) { cudaEventCreate(&rec.a); cudaEventCreate(&rec.b); cudaEventRecord(rec.a, cur()); }
    k_bucket_affine<<<(unsigned)((n + 127) / 128), 128, 0, cur()>>>(f, n);
    if (g_prof_on) { cudaEventRecord(rec.b, cur()); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back(rec); }
    g_launches++;
    return ck(cudaGetLastError(), "kernel launch");
}
int launch_horner_jac(const HornerJac& f, uint64_t n) {
    if (!n) return 0;
    if (n > 4096) return launch_occ(f, n);                 // many MSMs: throughput-bound, one thread each
    ProfRec rec{"HornerJacCoop", nullptr, nullptr, n * 32};
    if (g_prof_on) { cudaEventCreate(&rec.a); cudaEventCreate(&rec.b); cudaEventRecord(rec.a, cur()); }
    k_horner_jac_coop<<<(unsigned)n, 32, 0, cur()>>>(f, n);
    if (g_prof_on) { cudaEventRecord(rec.b, cur()); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back(rec); }
    g_launches++;
    return ck(cudaGetLastError(), "kernel launch");
}
void* scratch_alloc(size_t bytes);
void scratch_free(void* p);
// Horner pass over XYZZ window sums (the table-lookup MSMs of the prover).  Thousands of MSMs: one thread each.  A few
// MSMs (one proof per call): the 258 dependent doublings of a thread are 2.8 ms per round - the window sums are turned
// into Jacobian points and the warp-cooperative pass above takes over (1.2 ms).  Same group element either way.
int launch_horner_xyzz(const MsmShape& hs, const Xyzz* partial, Jac* out, uint64_t M) {
    if (M > 1024) return launch_occ(Horner{hs, partial, out}, M);
    Jac* tmp = (Jac*)scratch_alloc(M * hs.W * sizeof(Jac));
    if (!tmp) return fail("Horner: scratch allocation failed");
    int rc = launch(XyzzToJac{partial, tmp}, M * hs.W);
    if (!rc) rc = launch_horner_jac(HornerJac{hs.W, hs.c, tmp, out}, M);
    scratch_free(tmp);
    return rc;
}
void* scratch_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes ? bytes : 1, cur()) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void scratch_free(void* p) { if (p) cudaFreeAsync(p, cur()); }
#else
// ---- host emulation (test seam) ----
template <int BLOCK = 128, int MINB = 1, class F>
int launch(const F& f, uint64_t n) {
    g_launches++;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t t = 0; t < (int64_t)n; t++) f((uint64_t)t);
    return 0;
}
template <class F>
int launch_occ(const F& f, uint64_t n) { return launch(f, n); }
template <class F>
int launch_decomp(const F& f, uint64_t n) { return launch(f, n); }
int launch_sort_digits(const SortDigits& f, uint64_t n) { return launch(f, n); }
int launch_bucket_affine(const BucketAccumulateAffine& f, uint64_t n) { return launch(f, n); }
int launch_horner_jac(const HornerJac& f, uint64_t n) { return launch(f, n); }
int launch_horner_xyzz(const MsmShape& hs, const Xyzz* partial, Jac* out, uint64_t M) { return launch(Horner{hs, partial, out}, M); }
void* scratch_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void scratch_free(void* p) { free(p); }
#endif

int need_init() {
    if (g_device < 0) return fail("cpg_init has not been called (no CUDA device selected)");
    return 0;
}
#define NEED_INIT() do { if (int rc_ = need_init()) return rc_; } while (0)

struct Scratch {  // frees on scope exit (stream-ordered)
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) scratch_free(p); }
    template <class T> T* get(size_t count) {
        void* p = scratch_alloc(count * sizeof(T));
        if (p) ptrs.push_back(p);
        return (T*)p;
    }
};

uint32_t windows_for(uint32_t c) { return (256 + c - 1) / c; }

Recode make_recode(uint32_t c) {
    Recode rc; rc.c = c; rc.W = windows_for(c);
    for (int i = 0; i < 8; i++) rc.C[i] = 0;
    for (uint32_t w = 0; w + 1 < rc.W; w++) { uint32_t bit = c - 1 + c * w; rc.C[bit >> 5] |= 1u << (bit & 31); }
    return rc;
}

// modmul-count model of the bucket method: n*W mixed adds + 2*NB*W full adds + Horner
uint32_t pick_window(size_t n) {
    double best = 1e300; uint32_t bc = 4;
    for (uint32_t c = 3; c <= 16; c++) {
        double W = windows_for(c), NB = (double)(1u << (c - 1));
        double cost = (double)n * W * 10.0 + W * NB * 2.0 * 14.0 + W * (c * 9.0 + 14.0);
        if (cost < best) { best = cost; bc = c; }
    }
    return bc;
}

// chunk width of the level-wise window reduction (ReduceLevel); CPG_REDUCE_CH overrides it for tuning
uint32_t reduce_ch() {
    static uint32_t ch = [] {
        const char* e = getenv("CPG_REDUCE_CH");
        uint32_t v = e ? (uint32_t)atoi(e) : 4;
        return (v == 2 || v == 4 || v == 8 || v == 16) ? v : 4u;
    }();
    return ch;
}

// Single MSMs are launched with few (msm, window) pairs, so besides the total work the longest SERIAL
// chain matters: a bucket list is walked by one thread, and a window whose digit has only tw bits
// (the top window holds 255 - c(W-1) bits) concentrates n terms in 2^tw lists.  Estimated time =
// total products / pipe rate + serial products * single-thread product latency (0.9 us measured).
uint32_t pick_window_large(size_t n, size_t B = 1) {
    double best = 1e300; uint32_t bc = 8;
    const double ch = reduce_ch();
    for (uint32_t c = 4; c <= 16; c++) {
        double W = windows_for(c), NB = (double)(1u << (c - 1));
        int tw = 255 - (int)c * ((int)W - 1);
        double top_lists = tw <= 0 ? 1.0 : (double)(1u << tw);
        if (top_lists > NB) top_lists = NB;
        double lam = (double)n / NB, lam_top = (double)n / top_lists / (tw <= 0 ? 2.0 : 1.0);
        double longest = std::max(lam + 4.0 * std::sqrt(lam) + 4.0, lam_top + 4.0 * std::sqrt(lam_top));
        if (longest > (double)n) longest = (double)n;
        double levels = std::ceil((c - 1) / std::log2(ch));
        double serial = longest * 10.0 + levels * (2.0 * ch - 3.0) * 14.0 + levels * 14.0 + (c - 1) * 9.0;
        double total = (double)n * W * 10.0 + W * NB * 3.0 * 14.0;
        double t = (double)B * total / 2.7e10 + serial * 0.9e-6;
        if (t < best) { best = t; bc = c; }
    }
    return bc;
}

const uint8_t GEN48[48] = {0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                           0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                           0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};

struct FixedTable {
    FixedShape s;
    Aff* table;      // device
    size_t bytes;
};

}  // namespace

extern "C" {

const char* cpg_last_error(void) { return g_err.c_str(); }

#ifndef CPG_HOST_EMU
const char* cpg_backend(void) { return "cuda-sm_100a"; }
int cpg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
#else
const char* cpg_backend(void) { return "host-emulation-test-seam"; }
int cpg_device_count(void) { return 1; }
#endif

int cpg_init(int device) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_device >= 0) {
        if (g_device != device) return fail("cpg_init: already initialised on another device (one process per GPU)");
        return 0;
    }
#ifndef CPG_HOST_EMU
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail("cpg_init: no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail("cpg_init: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("cpg_init: kernels are built for sm_100a (B200) only");
    CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&g_ev0));
    CK(cudaEventCreate(&g_ev1));
    // keep freed scratch in the pool instead of returning it to the driver after every call
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ULL;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // A block freed on one stream may be handed to an allocation on another stream by making that stream WAIT for the
    // free (an "internal dependency").  For pipelines that run on several streams at once that wait is a hidden
    // serialisation.  CPG_POOL_INTERNAL_DEPS=0 forbids it (measured: no effect on the verifier's 24 streams, and the pool then
    // has to grow - allocation spikes of 10+ ms in the pipelined MSM experiment), so the driver's default stays.
    if (const char* e = getenv("CPG_POOL_INTERNAL_DEPS")) {
        int on = atoi(e) != 0;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &on));
    }
#endif
    g_device = device;
    // device-resident generator: decompress its wire encoding once
    uint8_t* d48 = (uint8_t*)cpg_malloc(48);
    Aff* daff = (Aff*)cpg_malloc(sizeof(Aff));
    uint8_t* derr = (uint8_t*)cpg_malloc(1);
    g_generator = (Jac*)cpg_malloc(sizeof(Jac));
    if (!d48 || !daff || !derr || !g_generator) { g_device = -1; return fail("cpg_init: allocation failed"); }
    int rc = cpg_h2d(d48, GEN48, 48);
    if (!rc) rc = cpg_g1_decompress(d48, 1, 0, daff, derr);
    if (!rc) rc = cpg_g1_aff_to_jac(daff, 1, g_generator);
    uint8_t e = 1;
    if (!rc) rc = cpg_d2h(&e, derr, 1);
    cpg_free(d48); cpg_free(daff); cpg_free(derr);
    if (rc || e) { g_device = -1; return fail("cpg_init: generator self-check failed: " + g_err); }
    return 0;
}

int cpg_set_stream(void* s) {
#ifndef CPG_HOST_EMU
    t_stream = (cudaStream_t)s; t_stream_set = (s != nullptr);
#else
    (void)s;
#endif
    return 0;
}
int cpg_sync(void) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaStreamSynchronize(cur()));
#endif
    return 0;
}
void* cpg_malloc(size_t bytes) {
    if (need_init()) return nullptr;
#ifndef CPG_HOST_EMU
    void* p = nullptr;
    if (ck(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
    return p;
#else
    return malloc(bytes ? bytes : 1);
#endif
}
int cpg_free(void* p) {
    if (!p) return 0;
#ifndef CPG_HOST_EMU
    CK(cudaFree(p));
#else
    free(p);
#endif
    return 0;
}
int cpg_memset(void* p, int v, size_t bytes) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaMemsetAsync(p, v, bytes, cur()));
#else
    memset(p, v, bytes);
#endif
    return 0;
}
int cpg_h2d(void* d, const void* h, size_t bytes) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, cur()));
#else
    memcpy(d, h, bytes);
#endif
    return 0;
}
int cpg_d2h(void* h, const void* d, size_t bytes) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, cur()));
    CK(cudaStreamSynchronize(cur()));
#else
    memcpy(h, d, bytes);
#endif
    return 0;
}
// device -> PINNED host memory on the current stream, no synchronisation (internal: results of one lane / sub-batch
// leave while the others still compute; the caller synchronises once)
static int d2h_async(void* h_pinned, const void* d, size_t bytes) {
#ifndef CPG_HOST_EMU
    CK(cudaMemcpyAsync(h_pinned, d, bytes, cudaMemcpyDeviceToHost, cur()));
#else
    memcpy(h_pinned, d, bytes);
#endif
    return 0;
}
int cpg_d2d(void* dst, const void* src, size_t bytes) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, cur()));
#else
    memmove(dst, src, bytes);
#endif
    return 0;
}
void* cpg_host_alloc(size_t bytes) {
#ifndef CPG_HOST_EMU
    void* p = nullptr;
    if (ck(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost")) return nullptr;
    return p;
#else
    return malloc(bytes ? bytes : 1);
#endif
}
int cpg_host_free(void* p) {
    if (!p) return 0;
#ifndef CPG_HOST_EMU
    CK(cudaFreeHost(p));
#else
    free(p);
#endif
    return 0;
}
int cpg_timer_start(void) {
    NEED_INIT();
#ifndef CPG_HOST_EMU
    CK(cudaEventRecord(g_ev0, cur()));
#endif
    return 0;
}
int cpg_timer_stop(float* ms) {
    NEED_INIT();
    *ms = 0.f;
#ifndef CPG_HOST_EMU
    CK(cudaEventRecord(g_ev1, cur()));
    CK(cudaEventSynchronize(g_ev1));
    CK(cudaEventElapsedTime(ms, g_ev0, g_ev1));
#endif
    return 0;
}
uint64_t cpg_launch_count(void) { return g_launches.load(); }

/* ---- per-kernel profile ---- */
int cpg_profile_enable(int on) {
#ifndef CPG_HOST_EMU
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
#else
    (void)on;
#endif
    return 0;
}
int cpg_profile_reset(void) {
#ifndef CPG_HOST_EMU
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    if (!g_d_counters) cudaMalloc(&g_d_counters, 2 * sizeof(unsigned long long));
    if (g_d_counters) cudaMemset(g_d_counters, 0, 2 * sizeof(unsigned long long));
#endif
    return 0;
}
int cpg_profile_report(char* buf, size_t cap) {
    if (!buf || !cap) return fail("cpg_profile_report: no buffer");
    buf[0] = 0;
#ifndef CPG_HOST_EMU
    NEED_INIT();
    CK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    struct Agg { const char* name; double ms; uint64_t launches, threads; };
    std::vector<Agg> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        size_t i = 0;
        for (; i < agg.size(); i++) if (!strcmp(agg[i].name, r.name)) break;
        if (i == agg.size()) agg.push_back(Agg{r.name, 0.0, 0, 0});
        agg[i].ms += ms; agg[i].launches++; agg[i].threads += r.threads;
    }
    // CPG_PROFILE_TRACE=<file>: also a timeline, one line per launch - name, start and end in ms since the first recorded
    // launch (event timestamps are comparable across streams): shows what overlaps what
    if (const char* path = getenv("CPG_PROFILE_TRACE")) {
        if (FILE* tf = fopen(path, "w")) {
            for (auto& r : g_prof) {
                float t0 = 0.f, t1 = 0.f;
                if (cudaEventElapsedTime(&t0, g_prof[0].a, r.a) != cudaSuccess || cudaEventElapsedTime(&t1, g_prof[0].a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
                fprintf(tf, "%-22s %10.4f %10.4f  %llu threads\n", r.name, t0, t1, (unsigned long long)r.threads);
            }
            fclose(tf);
        }
    }
    std::string out = "{";
    if (g_d_counters) {
        unsigned long long c[2] = {0, 0};
        cudaMemcpy(c, g_d_counters, sizeof c, cudaMemcpyDeviceToHost);
        char line[160];
        snprintf(line, sizeof line, "\"_counters\": {\"fixed_msm_terms\": %llu, \"var_table_msm_terms\": %llu}%s", c[0], c[1], agg.empty() ? "" : ", ");
        out += line;
    }
    for (size_t i = 0; i < agg.size(); i++) {
        char line[256];
        snprintf(line, sizeof line, "%s\"%s\": {\"ms\": %.6f, \"launches\": %llu, \"threads\": %llu}", i ? ", " : "",
                 agg[i].name, agg[i].ms, (unsigned long long)agg[i].launches, (unsigned long long)agg[i].threads);
        out += line;
    }
    out += "}";
    if (out.size() + 1 > cap) return fail("cpg_profile_report: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
#else
    if (cap >= 3) { buf[0] = '{'; buf[1] = '}'; buf[2] = 0; }
#endif
    return 0;
}

/* ---- serialisation ---- */
int cpg_g1_decompress(const uint8_t* d_in, size_t k, int check, void* d_out, uint8_t* d_err) {
    NEED_INIT();
    if (check) return launch_occ(DecompressChecked{d_in, (Aff*)d_out, d_err}, k);
    return launch_decomp(Decompress{d_in, (Aff*)d_out, d_err}, k);
}
int cpg_g1_compress(const void* d_jac, size_t k, uint8_t* d_out) {
    NEED_INIT();
    return launch_occ(CompressJac{(const Jac*)d_jac, d_out}, k);
}
int cpg_g1_compress_aff(const void* d_aff, size_t k, uint8_t* d_out) {
    NEED_INIT();
    return launch(CompressAff{(const Aff*)d_aff, d_out}, k);
}
int cpg_g1_aff_to_jac(const void* d_aff, size_t k, void* d_out) {
    NEED_INIT();
    return launch(AffToJac{(const Aff*)d_aff, (Jac*)d_out}, k);
}
int cpg_g1_jac_to_aff(const void* d_jac, size_t k, void* d_out) {
    NEED_INIT();
    return launch_occ(JacToAff{(const Jac*)d_jac, (Aff*)d_out}, k);
}
int cpg_g1_generator(void* d_out) {
    NEED_INIT();
    return cpg_d2d(d_out, g_generator, sizeof(Jac));
}
int cpg_g1_identity(void* d_out) {
    NEED_INIT();
    Jac h;
    for (int i = 0; i < 12; i++) { h.X.l[i] = H_FQ_R[i]; h.Y.l[i] = H_FQ_R[i]; h.Z.l[i] = 0; }
#ifndef CPG_HOST_EMU
    // pageable source: the copy is staged before the call returns
    CK(cudaMemcpyAsync(d_out, &h, sizeof h, cudaMemcpyHostToDevice, cur()));
    CK(cudaStreamSynchronize(cur()));
    return 0;
#else
    memcpy(d_out, &h, sizeof h);
    return 0;
#endif
}

/* ---- element-wise group law ---- */
int cpg_g1_add(const void* a, const void* b, size_t k, void* out) {
    NEED_INIT();
    return launch(AddPoints{(const Jac*)a, (const Jac*)b, (Jac*)out, 0}, k);
}
int cpg_g1_sub(const void* a, const void* b, size_t k, void* out) {
    NEED_INIT();
    return launch(AddPoints{(const Jac*)a, (const Jac*)b, (Jac*)out, 1}, k);
}
int cpg_g1_neg(const void* a, size_t k, void* out) {
    NEED_INIT();
    return launch(NegPoints{(const Jac*)a, (Jac*)out}, k);
}
int cpg_g1_eq(const void* a, const void* b, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(EqPoints{(const Jac*)a, (const Jac*)b, out}, k);
}
int cpg_g1_is_identity(const void* a, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(IsInfPoints{(const Jac*)a, out}, k);
}
int cpg_g1_mul(const void* p, const uint8_t* scalars, size_t k, size_t group, void* out) {
    NEED_INIT();
    if (!group) return fail("cpg_g1_mul: group must be >= 1");
    return launch(MulPoints{(const Jac*)p, (const uint32_t*)scalars, group, (Jac*)out}, k);
}
int cpg_g1_fold(const void* L, const void* R, const uint8_t* x, size_t rows, size_t m, void* out) {
    NEED_INIT();
    if (!m) return 0;
    return launch(FoldPoints{(const Jac*)L, (const Jac*)R, (const uint32_t*)x, m, (Jac*)out}, rows * m);
}

/* ---- batched Pippenger ---- */
static int msm_batched_impl(const void* d_bases, size_t base_stride, const uint32_t* d_base_off, const uint8_t* d_scalars,
                            size_t B, size_t n, int window, void* d_out, uint32_t w0 = 0, uint32_t wn = 0, size_t sc_stride = 0, size_t sc_off = 0);
int cpg_g1_msm_batched(const void* d_bases, size_t base_stride, const uint8_t* d_scalars,
                       size_t B, size_t n, int window, void* d_out) {
    return msm_batched_impl(d_bases, base_stride, nullptr, d_scalars, B, n, window, d_out);
}
int cpg_g1_msm_batched_off(const void* d_bases, const uint32_t* d_base_off, const uint8_t* d_scalars,
                           size_t B, size_t n, int window, void* d_out) {
    if (!d_base_off) return fail("cpg_g1_msm_batched_off: null offsets");
    return msm_batched_impl(d_bases, 0, d_base_off, d_scalars, B, n, window, d_out);
}
int cpg_msm_window_count(size_t n, int window) {
    uint32_t c = window > 0 ? (uint32_t)window : pick_window_large(n ? n : 1);
    return (int)windows_for(c);
}
/* window width cpg_g1_msm_batched picks (window = 0) for B MSMs of n terms each */
int cpg_msm_pick_window_batched(size_t B, size_t n) {
    if (!n) n = 1;
    if (!B) B = 1;
    const bool few = g_msm_path == 0 ? B * 32 < 8192 : g_msm_path == 2;
    return (int)((few || n > 2048) ? pick_window_large(n, B) : pick_window(n));
}
int cpg_msm_set_accumulate(int mode) {
    if (mode < 0 || mode > 1) return fail("cpg_msm_set_accumulate: 0 (mixed XYZZ additions) or 1 (batched affine additions)");
    g_msm_affine = mode;
    return 0;
}
int cpg_msm_force_path(int path) {
    if (path < 0 || path > 2) return fail("cpg_msm_force_path: 0 (by shape), 1 (per-window threads) or 2 (per-term threads)");
    g_msm_path = path;
    return 0;
}
/* window width this library uses for ONE n-term MSM (a function of n only, so all ranks agree) */
int cpg_msm_pick_window(size_t n) { return (int)pick_window_large(n ? n : 1); }
/* window sums S_w, w in [w_begin, w_end), of ONE n-term MSM as Jacobian points (the unit of the
 * multi-GPU window split: every rank computes a slice, the slices are all-gathered, then combined) */
int cpg_g1_msm_window_sums(const void* d_bases, const uint8_t* d_scalars, size_t n, int window,
                           int w_begin, int w_end, void* d_out_jac) {
    if (window <= 0) return fail("cpg_g1_msm_window_sums: the window width must be given (all ranks must agree)");
    int W = (int)windows_for((uint32_t)window);
    if (w_begin < 0 || w_end > W || w_begin >= w_end) return fail("cpg_g1_msm_window_sums: bad window range");
    if (!n) { for (int w = w_begin; w < w_end; w++) if (int rc = cpg_g1_identity((Jac*)d_out_jac + (w - w_begin))) return rc; return 0; }
    return msm_batched_impl(d_bases, 0, nullptr, d_scalars, 1, n, window, d_out_jac, (uint32_t)w_begin, (uint32_t)(w_end - w_begin));
}
/* out = sum_w 2^(c w) S_w over all W = cpg_msm_window_count windows (Horner) */
int cpg_g1_msm_combine_windows(const void* d_wsums_jac, int window, void* d_out_jac) {
    NEED_INIT();
    if (window <= 0) return fail("cpg_g1_msm_combine_windows: bad window");
    return launch_horner_jac(HornerJac{windows_for((uint32_t)window), (uint32_t)window, (const Jac*)d_wsums_jac, (Jac*)d_out_jac}, 1);
}
#ifndef CPG_HOST_EMU
// ONE large MSM, pipelined over slices of its windows (top slice first) - EXPERIMENT, off by default (CPG_MSM_SLICES > 1
// turns it on).  The tail of the plain pipeline is pure latency: 8 reduction levels + the 255 dependent doublings of the
// Horner pass = 2.6 of the 9.3 ms of an n = 2^20 MSM, during which the GPU is nearly idle.  Here the GPU-filling stages
// of all slices (digit sort, bucket accumulation) run back to back on one stream, and each slice's level-wise
// reduction + its segment of the Horner pass run on a second, high-priority stream under the next slice's
// accumulation.  Results are bit-identical (same window sums, same Horner order) - and it is NOT faster
// (profiles/r02_ab_msm_pipeline.txt, r02_msm_trace_slices4.txt): the overlap happens, but a latency chain that shares
// the SMs with 12 accumulating warps each runs 2-3x slower (a reduction level 0.10 -> 0.2-0.7 ms, a Horner segment
// 0.32 -> 0.87 ms), the per-slice sort costs 0.44 ms four times instead of 0.94 ms once, and the sliced accumulations
// add up to 6.4 instead of 5.7 ms: n = 2^20 9.36 ms plain, 9.2 ms in 2 slices, 9.8-10.8 in 4, 13-15 in 8.
int msm_pipe_slices() { static const int v = getenv("CPG_MSM_SLICES") ? atoi(getenv("CPG_MSM_SLICES")) : 1; return v; }
size_t msm_pipe_min_n() { static const size_t v = getenv("CPG_MSM_PIPE_MIN_N") ? (size_t)atoll(getenv("CPG_MSM_PIPE_MIN_N")) : ((size_t)1 << 18); return v; }
static int msm_large_pipelined(const void* d_bases, const uint8_t* d_scalars, size_t n, uint32_t c, void* d_out, int slices) {
    const uint32_t W = windows_for(c), per = (W + (uint32_t)slices - 1) / (uint32_t)slices;
    thread_local cudaStream_t st_low = nullptr, st_high = nullptr;
    if (!st_low) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));                       // lo = least, hi = greatest priority
        CK(cudaStreamCreateWithPriority(&st_low, cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&st_high, cudaStreamNonBlocking, hi));
    }
    const cudaStream_t main_stream = cur();
    const bool saved_set = t_stream_set; const cudaStream_t saved = t_stream;
    Jac* wsum = (Jac*)scratch_alloc((size_t)W * sizeof(Jac));
    if (!wsum) return fail("cpg_g1_msm_batched: scratch allocation failed");
    std::vector<cudaEvent_t> evs;
    auto new_event = [&]() { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); return e; };
    cudaEvent_t fork = new_event();
    int rc = ck(cudaEventRecord(fork, main_stream), "cudaEventRecord");
    if (!rc) rc = ck(cudaStreamWaitEvent(st_low, fork, 0), "cudaStreamWaitEvent");
    if (!rc) rc = ck(cudaStreamWaitEvent(st_high, fork, 0), "cudaStreamWaitEvent");
    int k = 0;
    for (uint32_t w_end = W; w_end > 0 && !rc; k++) {
        const uint32_t w_begin = w_end > per ? w_end - per : 0;
        t_stream = st_low; t_stream_set = true;
        t_pipe_signal = new_event(); t_pipe_tail = st_high;
        rc = msm_batched_impl(d_bases, 0, nullptr, d_scalars, 1, n, (int)c, wsum + w_begin, w_begin, w_end - w_begin);
        t_pipe_signal = nullptr; t_pipe_tail = nullptr;
        t_stream = st_high; t_stream_set = true;                             // (the call left it there after its accumulation)
        if (!rc) { HornerJac hj{W, c, wsum, (Jac*)d_out}; hj.w_lo = w_begin; hj.w_hi = w_end; hj.cont = k > 0 ? 1u : 0u; rc = launch_horner_jac(hj, 1); }
        w_end = w_begin;
    }
    // join: the caller's stream continues after both
    cudaEvent_t e_low = new_event(), e_high = new_event();
    if (cudaEventRecord(e_low, st_low) == cudaSuccess) cudaStreamWaitEvent(main_stream, e_low, 0);
    if (cudaEventRecord(e_high, st_high) == cudaSuccess) cudaStreamWaitEvent(main_stream, e_high, 0);
    t_stream = saved; t_stream_set = saved_set;
    scratch_free(wsum);
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    return rc;
}
#endif
// wn = 0: all windows and the final Horner; wn > 0: only windows [w0, w0+wn), output = their sums (B must be 1)
// sc_stride / sc_off: the scalar of (msm m, term i) is d_scalars[m*sc_stride + sc_off + i] (sc_stride = 0: rows of n, no offset)
static int msm_batched_impl(const void* d_bases, size_t base_stride, const uint32_t* d_base_off, const uint8_t* d_scalars,
                            size_t B, size_t n, int window, void* d_out, uint32_t w0, uint32_t wn, size_t sc_stride, size_t sc_off) {
    NEED_INIT();
    if (!B) return 0;
    if (n == 0) {  // empty sums are the identity (compute_MSM returns Z1 for empty input)
        for (size_t b = 0; b < B; b++) if (int rc = cpg_g1_identity((Jac*)d_out + b)) return rc;
        return 0;
    }
    if (n >= 0x7fffffffULL) return fail("cpg_g1_msm_batched: n too large");
    // thousands of small MSMs: per-(msm, window) threads; few (or big) MSMs: per-term threads with atomics,
    // buckets ordered by list length, level-wise window reduction
    const bool few = g_msm_path == 0 ? B * 32 < 8192 : g_msm_path == 2;
    uint32_t c = window > 0 ? (uint32_t)window : (uint32_t)cpg_msm_pick_window_batched(B, n);
    if (c < 2 || c > 16) return fail("cpg_g1_msm_batched: window must be in [2, 16]");
    Recode rc = make_recode(c);
    const bool slice = wn != 0;
    if (slice && B != 1) return fail("cpg_g1_msm_batched: window slices are for single MSMs");
    MsmShape s; s.n = (uint32_t)n; s.c = c; s.W = rc.W; s.NB = 1u << (c - 1); s.base_stride = base_stride; s.base_off = nullptr;
    s.w0 = slice ? w0 : 0; s.wn = slice ? wn : rc.W;
    s.sc_stride = sc_stride ? sc_stride : n; s.sc_off = sc_off;
    const bool large = few || n > 2048 || s.NB > 256;
    if (g_msm_path == 1 && large) return fail("cpg_g1_msm_batched: the per-window path handles n <= 2048 and windows <= 9 bits");
#ifndef CPG_HOST_EMU
    if (B == 1 && !slice && large && !d_base_off && !sc_stride && !sc_off && n >= msm_pipe_min_n() && msm_pipe_slices() > 1 && rc.W >= 2u * (uint32_t)msm_pipe_slices())
        return msm_large_pipelined(d_bases, d_scalars, n, c, d_out, msm_pipe_slices());
#endif
    // reduction levels: NB = prod ch_j
    uint32_t nlev = 0; uint8_t lg_ch[16]; uint32_t chs[16];
    size_t lvl_elems = 0;                                          // largest level output, in points per (msm, window)
    if (large) {
        for (uint32_t len = s.NB; len > 1;) {
            uint32_t ch = reduce_ch(); while (ch > len) ch >>= 1;
            chs[nlev] = ch; uint8_t lg = 0; while ((1u << lg) < ch) lg++; lg_ch[nlev] = lg;
            len /= ch; nlev++;
            lvl_elems = std::max(lvl_elems, (size_t)(nlev + 1) * len);
        }
    }
    // bound scratch to ~6 GiB per chunk of MSMs
    size_t per_msm = (size_t)s.wn * ((size_t)(s.NB + 1) * 8 + (size_t)n * 4 + (size_t)s.NB * (sizeof(Xyzz) + 6) + 2 * sizeof(Xyzz) * (lvl_elems + 1)) + (size_t)s.W * n * 2;
    // batched affine accumulation: two scratch arrays of `cap` points per thread.  Per-window path: a thread owns a whole
    // window, cap = (n + NB) / 2 always suffices.  Per-term path: a thread owns bpt buckets holding ~512 points at the
    // density of the densest window (the top window has only tw bits: 2^tw non-empty buckets).
    const bool affine = g_msm_affine != 0;
    uint32_t bpt = s.NB, cap = (uint32_t)((n + s.NB) / 2 + 1);
    if (affine && large) {
        const int tw = 255 - (int)c * ((int)rc.W - 1);
        const double top_buckets = tw >= (int)c - 1 ? (double)s.NB : (double)(1u << (tw > 0 ? tw : 0));
        const double dens = (double)n / (slice && !(w0 + wn == rc.W) ? (double)s.NB : std::min((double)s.NB, top_buckets));
        bpt = 1; while (bpt < s.NB && (double)(2 * bpt) * dens <= 512.0) bpt *= 2;
        // ... but not so many that the launch cannot fill the GPU (two waves of 148 SMs x 3 blocks x 128 threads),
        // as long as a thread keeps ~128 points (fewer, and the shared inversions stop paying)
        while (bpt > 1 && (double)B * s.wn * (s.NB / bpt) < 2.0 * 148 * 3 * 128 && (double)(bpt / 2) * dens >= 128.0) bpt /= 2;
        const double pts = (double)bpt * dens;
        cap = (uint32_t)((pts + 6.0 * std::sqrt(pts) + 32.0 + bpt) / 2.0) + 1;
    }
    if (affine) per_msm += (size_t)s.wn * (s.NB / bpt) * cap * 2 * sizeof(Aff);
    size_t chunk = (size_t)6 << 30;
    chunk = chunk / per_msm; if (chunk < 1) chunk = 1; if (chunk > B) chunk = B;
    if (large) while (chunk > 1 && (uint64_t)chunk * s.wn * s.NB >= 0xffffffffULL) chunk /= 2;
    for (size_t b0 = 0; b0 < B; b0 += chunk) {
        size_t nb = B - b0 < chunk ? B - b0 : chunk;
        s.B = (uint32_t)nb;
        uint64_t BW = (uint64_t)nb * s.wn;
        if (large && BW * s.NB >= 0xffffffffULL) return fail("cpg_g1_msm_batched: window too wide for this many windows");
        Scratch sc;
        uint32_t* boff = sc.get<uint32_t>(BW * (s.NB + 1));
        uint32_t* sorted = sc.get<uint32_t>(BW * n);
        const bool balanced = !large && !affine;                    // length ranks: only the XYZZ accumulation (thread = bucket) uses them
        uint16_t* rank = balanced ? sc.get<uint16_t>(BW * s.NB) : nullptr;
        if (balanced && !rank) return fail("cpg_g1_msm_batched: scratch allocation failed");
        Xyzz* buckets = sc.get<Xyzz>(BW * s.NB);
        int16_t* dig = sc.get<int16_t>((uint64_t)nb * n * s.W);
        if (!boff || !sorted || !buckets || !dig) return fail("cpg_g1_msm_batched: scratch allocation failed");
        const uint32_t* ks = (const uint32_t*)d_scalars + (uint64_t)b0 * s.sc_stride * 8;
        const Aff* bases = (const Aff*)d_bases + (d_base_off ? 0 : (uint64_t)b0 * base_stride);
        s.base_off = d_base_off ? d_base_off + b0 : nullptr;
        if (int r = launch(RecodeDigits{s, rc, ks, dig}, (uint64_t)nb * n)) return r;
        if (!large) {
            Xyzz* wsum = sc.get<Xyzz>(BW);
            if (!wsum) return fail("cpg_g1_msm_batched: scratch allocation failed");
            if (int r = launch_sort_digits(SortDigits{s, dig, boff, sorted, rank}, BW)) return r;
            if (affine) {
                const uint64_t T = (((uint64_t)nb + 31) / 32) * 32 * s.wn;
                Aff* bufA = sc.get<Aff>(T * cap); Aff* bufB = sc.get<Aff>(T * cap);
                if (!bufA || !bufB) return fail("cpg_g1_msm_batched: scratch allocation failed");
                if (int r = launch_bucket_affine(BucketAccumulateAffine{s, bases, boff, sorted, BW, s.NB, 1u, cap, T, bufA, bufB, buckets}, T)) return r;
            } else {
                uint64_t nthreads = (((uint64_t)nb + 31) / 32) * 32 * s.wn * s.NB;
                if (int r = launch<128, 3>(BucketAccumulate{s, bases, boff, sorted, rank, nullptr, BW, buckets}, nthreads)) return r;
            }
            if (int r = launch_occ(WindowReduce{s, buckets, wsum}, BW)) return r;
            if (int r = launch_occ(Horner{s, wsum, (Jac*)d_out + b0}, nb)) return r;
            continue;
        }
        // counting sort of the digits: counts (atomics), three-pass scan, scatter
        const uint32_t sch = s.NB < SCAN_CH ? s.NB : SCAN_CH, nsc = s.NB / sch;
        uint32_t* cnt = sc.get<uint32_t>(BW * (s.NB + 1));
        uint32_t* ctot = sc.get<uint32_t>(BW * nsc);
        uint32_t* hist = sc.get<uint32_t>(LEN_BINS);
        uint32_t* order = sc.get<uint32_t>(BW * s.NB);
        Xyzz* lvA = sc.get<Xyzz>(BW * lvl_elems);
        Xyzz* lvB = sc.get<Xyzz>(BW * lvl_elems);
        Jac* wsum = slice ? (Jac*)d_out : sc.get<Jac>(BW);
        if (!cnt || !ctot || !hist || !order || !lvA || !lvB || !wsum) return fail("cpg_g1_msm_batched: scratch allocation failed");
        if (int r = cpg_memset(cnt, 0, BW * (s.NB + 1) * 4)) return r;
        if (int r = cpg_memset(hist, 0, LEN_BINS * 4)) return r;
        if (int r = launch(LargeCount{s, dig, cnt}, (uint64_t)nb * n)) return r;
        if (int r = launch(LargeScanChunks{s, sch, nsc, cnt, ctot}, BW * nsc)) return r;
        if (int r = launch(LargeScanTop{nsc, ctot}, BW)) return r;
        if (int r = launch(LargeScanApply{s, sch, nsc, cnt, ctot, boff}, BW * nsc)) return r;
        if (int r = launch(LargeScatter{s, dig, boff, cnt, sorted}, (uint64_t)nb * n)) return r;
        if (affine) {
            const uint64_t T = BW * (s.NB / bpt);
            Aff* bufA = sc.get<Aff>(T * cap); Aff* bufB = sc.get<Aff>(T * cap);
            if (!bufA || !bufB) return fail("cpg_g1_msm_batched: scratch allocation failed");
            if (int r = launch_bucket_affine(BucketAccumulateAffine{s, bases, boff, sorted, BW, bpt, 0u, cap, T, bufA, bufB, buckets}, T)) return r;
        } else {
            // buckets by list length, longest first
            if (int r = launch(LenHist{s, boff, hist}, BW * s.NB)) return r;
            if (int r = launch(LenScan{hist}, 1)) return r;
            if (int r = launch(LenScatter{s, boff, hist, order}, BW * s.NB)) return r;
            if (int r = launch<128, 3>(BucketAccumulate{s, bases, boff, sorted, nullptr, order, BW, buckets}, BW * s.NB)) return r;
        }
#ifndef CPG_HOST_EMU
        if (t_pipe_signal && t_pipe_tail) {
            // pipelined single MSM: everything after the accumulation - and the stream-ordered release of this call's
            // scratch - continues on the high-priority tail stream; the caller's stream is free for the next slice
            CK(cudaEventRecord(t_pipe_signal, cur()));
            CK(cudaStreamWaitEvent(t_pipe_tail, t_pipe_signal, 0));
            t_stream = t_pipe_tail; t_stream_set = true;
        }
#endif
        // level-wise window reduction
        const Xyzz* in = buckets; Xyzz* out = lvA;
        uint32_t len = s.NB;
        for (uint32_t j = 0; j < nlev; j++) {
            uint32_t ch = chs[j];
            if (int r = launch_occ(ReduceLevel{j + 1, ch, len, BW, in, out}, (uint64_t)(j + 1) * BW * (len / ch))) return r;
            len /= ch; in = out; out = (out == lvA) ? lvB : lvA;
        }
        ReduceFinal fin; fin.levels = nlev; fin.BW = BW; fin.in = in; fin.wsum = wsum;
        for (int k = 0; k < 16; k++) fin.lg_ch[k] = k < (int)nlev ? lg_ch[k] : 0;
        if (int r = launch(fin, BW)) return r;
        if (!slice) {
            if (int r = launch_horner_jac(HornerJac{s.W, s.c, wsum, (Jac*)d_out + b0}, nb)) return r;
        }
    }
    return 0;
}

/* ---- fixed-base tables ---- */
void* cpg_fixed_table_create(const void* d_bases, size_t nb, int window) {
    if (need_init()) return nullptr;
    if (!nb) { fail("cpg_fixed_table_create: empty base vector"); return nullptr; }
    uint32_t c = window > 0 ? (uint32_t)window : 8;
    if (c < 2 || c > 16) { fail("cpg_fixed_table_create: window must be in [2, 16]"); return nullptr; }
    FixedTable* t = new FixedTable;
    t->s.nb = (uint32_t)nb; t->s.c = c; t->s.W = windows_for(c); t->s.NB = 1u << (c - 1);
    uint64_t entries = (uint64_t)nb * t->s.W * t->s.NB;
    t->bytes = entries * sizeof(Aff);
    t->table = (Aff*)cpg_malloc(t->bytes);
    const uint64_t rows = (uint64_t)nb * t->s.W;
    const uint32_t S = t->s.NB < 64 ? t->s.NB : 64, nseg = t->s.NB / S;
    const uint64_t nthreads = rows * nseg;
    const uint64_t chunk = std::min<uint64_t>(nthreads, (uint64_t)148 * 3 * 128 * 4);       // threads per launch: bounds the scratch (96 B per entry in flight)
    Jac* hj = (Jac*)cpg_malloc(rows * sizeof(Jac));
    Fq* hz = (Fq*)cpg_malloc(rows * sizeof(Fq));
    Aff* heads = (Aff*)cpg_malloc(rows * sizeof(Aff));
    Fq* zs = (Fq*)cpg_malloc(chunk * S * sizeof(Fq));
    Fq* pz = (Fq*)cpg_malloc(chunk * S * sizeof(Fq));
    int rc = (!t->table || !hj || !hz || !heads || !zs || !pz) ? fail("cpg_fixed_table_create: allocation failed") : 0;
    if (!rc) rc = launch(FixedTableHeads{t->s, (const Aff*)d_bases, hj, hz, heads}, nb);
    for (uint64_t t0 = 0; t0 < nthreads && !rc; t0 += chunk)
        rc = launch_occ(FixedTableSegs{t->s, S, nseg, heads, t0, zs, pz, t->table}, std::min(chunk, nthreads - t0));
    if (!rc) rc = cpg_sync();
    cpg_free(hj); cpg_free(hz); cpg_free(heads); cpg_free(zs); cpg_free(pz);
    if (rc) { cpg_free(t->table); delete t; return nullptr; }
    return t;
}
int cpg_fixed_table_free(void* table) {
    if (!table) return 0;
    FixedTable* t = (FixedTable*)table;
    cpg_free(t->table);
    delete t;
    return 0;
}
size_t cpg_fixed_table_bytes(const void* table) { return table ? ((const FixedTable*)table)->bytes : 0; }

}  // extern "C"
// row_stride / row_off: the table's bases take entries row_off .. row_off + nb of coefficient rows that are row_stride
// entries long (a rank of a sharded proof holds the table of its own bases only); row_stride = 0: rows of nb, no offset
static int msm_fixed_impl(const void* table, const uint8_t* d_scalars, size_t B, int accumulate, void* d_out, size_t row_stride, size_t row_off);
extern "C" {
int cpg_g1_msm_fixed_batched(const void* table, const uint8_t* d_scalars, size_t B, int accumulate, void* d_out) {
    return msm_fixed_impl(table, d_scalars, B, accumulate, d_out, 0, 0);
}
}  // extern "C"
static int msm_fixed_impl(const void* table, const uint8_t* d_scalars, size_t B, int accumulate, void* d_out, size_t row_stride, size_t row_off) {
    NEED_INIT();
    if (!table) return fail("cpg_g1_msm_fixed_batched: null table");
    if (!B) return 0;
    const FixedTable* t = (const FixedTable*)table;
    if (!row_stride) row_stride = t->s.nb;
    Recode rc = make_recode(t->s.c);
    Scratch sc;
    // few MSMs over many bases (a single large proof): <= 64 bases per thread, then a tree over the partial sums
    // (fan-in 8) - every stage is a short chain instead of one thread walking hundreds of bases / partials
    uint32_t nchunk = 1;
    // (Whisk-size tables, ~131 bases: chunks of 16 - a chunk of 44 was a 0.43 ms chain in each of the 21 prover rounds)
    static const uint32_t big_chunk = getenv("CPG_FIXED_CHUNK") ? (uint32_t)atoi(getenv("CPG_FIXED_CHUNK")) : 64;   // bases per thread for tables of >= 1024 bases
    if ((uint64_t)B * t->s.W < 8192 && t->s.nb >= 128) nchunk = t->s.nb >= 1024 ? (t->s.nb + big_chunk - 1) / big_chunk : (t->s.nb + 15) / 16;
    Xyzz* partial = sc.get<Xyzz>((uint64_t)B * t->s.W * nchunk);
    if (!partial) return fail("cpg_g1_msm_fixed_batched: scratch allocation failed");
#ifndef CPG_HOST_EMU
    if (g_prof_on && g_d_counters) if (int r = launch(CountNonZero{(const uint32_t*)d_scalars, g_d_counters, t->s.nb, row_stride, row_off}, (uint64_t)B * t->s.nb)) return r;
#endif
    const uint64_t nthreads = nchunk > 1 ? (uint64_t)B * t->s.W * nchunk : (((uint64_t)B + 31) / 32) * 32 * t->s.W;
    if (int r = launch<128, 3>(FixedMsmWindow{t->s, rc, (uint32_t)B, nchunk, t->table, (const uint32_t*)d_scalars, row_stride, row_off, partial}, nthreads)) return r;
    uint32_t np = t->s.W * nchunk;
    const bool few = (uint64_t)B * t->s.W < 8192;           // a few MSMs: every stage a short chain (fan-in 8), also for small tables
    while (np > (few ? 8u : 64u)) {
        const uint32_t per = few ? 8 : 16, np_out = (np + per - 1) / per;
        Xyzz* stage = sc.get<Xyzz>((uint64_t)B * np_out);
        if (!stage) return fail("cpg_g1_msm_fixed_batched: scratch allocation failed");
        if (int r = launch_occ(SumPartialsRagged{np, per, np_out, partial, stage}, (uint64_t)B * np_out)) return r;
        partial = stage; np = np_out;
    }
    return launch_occ(SumWindows{np, partial, (Jac*)d_out, accumulate}, B);
}
extern "C" {

/* ---- Fr vectors ---- */
int cpg_fr_add(const uint8_t* a, const uint8_t* b, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(FrBinary{(const uint32_t*)a, (const uint32_t*)b, (uint32_t*)out, 0}, k);
}
int cpg_fr_sub(const uint8_t* a, const uint8_t* b, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(FrBinary{(const uint32_t*)a, (const uint32_t*)b, (uint32_t*)out, 1}, k);
}
int cpg_fr_mul(const uint8_t* a, const uint8_t* b, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(FrBinary{(const uint32_t*)a, (const uint32_t*)b, (uint32_t*)out, 2}, k);
}
int cpg_fr_inverse(const uint8_t* a, size_t k, uint8_t* out) {
    NEED_INIT();
    return launch(FrInverse{(const uint32_t*)a, (uint32_t*)out}, k);
}

/* ---- integer-pipe microbenchmark ---- */
#ifndef CPG_HOST_EMU
}  // extern "C"
namespace {
// 8 independent chains per thread.  The multiplicand is the accumulator's own low word, so nothing
// is loop-invariant (an invariant product gets strength-reduced to adds by ptxas - an earlier
// version of this probe measured the ALU pipe that way).  SASS: one IMAD.WIDE.U32 per step.
#define CPG_WIDE_DEP(lo, hi, a) asm volatile("{ .reg .u32 t; mov.u32 t, %0; mad.lo.cc.u32 %0, %2, t, %0; madc.hi.u32 %1, %2, t, %1; }" : "+r"(lo), "+r"(hi) : "r"(a))
__global__ void __launch_bounds__(256) k_imad_wide(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    // one fixed multiplier register, data-dependent multiplicand: the highest sustained rate of the
    // probe family in profiles/r01_imad_probe.txt (9.16 TMAC/s; distinct multipliers per chain: 8.5)
    uint32_t a = a0 + threadIdx.x, lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = j + threadIdx.x + b0; hi[j] = j + blockIdx.x; }
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) CPG_WIDE_DEP(lo[j], hi[j], a);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
    if (s == 0x12345678u) sink[0] = s;
}
// 32-bit IMAD (low half only), data-dependent: the full-rate integer multiply, for context
__global__ void __launch_bounds__(256) k_imad_lohi(uint64_t iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
    uint32_t a = a0 + threadIdx.x, x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = j + threadIdx.x + b0 + blockIdx.x;
    for (uint64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %1, %0, %0;" : "+r"(x[j]) : "r"(a));
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    if (s == 0x12345678u) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_fq_mul_chain(uint64_t iters, uint64_t* sink) {
    Fq x = Fq::one(), y = Fq::one();
    x.l[0] += threadIdx.x; y.l[1] += blockIdx.x;
    Fq u = y, v = x;
    for (uint64_t i = 0; i < iters; i++) { x = mul(x, y); u = mul(u, v); }
    if ((x.l[0] ^ u.l[3]) == 0x12345678u) sink[0] = x.l[1];
}
}  // namespace
extern "C" {
int cpg_bench_int_pipe(int kind, uint64_t iters, double* per_second, float* ms) {
    NEED_INIT();
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    int blocks = prop.multiProcessorCount * 8, threads = 256;
    uint64_t* sink = (uint64_t*)cpg_malloc(8);
    if (!sink) return 1;
    for (int rep = 0; rep < 2; rep++) {  // first pass warms up
        CK(cudaEventRecord(g_ev0, cur()));
        if (kind == 0) k_imad_wide<<<blocks, threads, 0, cur()>>>(iters, 12345u, 6789u, sink);
        else if (kind == 1) k_imad_lohi<<<blocks, threads, 0, cur()>>>(iters, 12345u, 6789u, sink);
        else k_fq_mul_chain<<<blocks, threads, 0, cur()>>>(iters, sink);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(g_ev1, cur()));
        CK(cudaEventSynchronize(g_ev1));
        CK(cudaEventElapsedTime(ms, g_ev0, g_ev1));
    }
    double ops = (double)blocks * threads * (double)iters * (kind == 2 ? 2.0 : 8.0);
    *per_second = ops / (*ms * 1e-3);
    cpg_free(sink);
    return 0;
}
#else
int cpg_bench_int_pipe(int, uint64_t, double* per_second, float* ms) {
    *per_second = 0; *ms = 0;
    return fail("cpg_bench_int_pipe: needs the CUDA build");
}
#endif

}  // extern "C"

#include "comm.inl"
#include "verify.inl"
#include "prove.inl"
#include "pyrandom.inl"

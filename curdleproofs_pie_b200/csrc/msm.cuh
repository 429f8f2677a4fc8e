// Kernel bodies (one functor = one kernel, one call = one thread) for the batched G1 hot path.
//
// Reference sites replaced (all under /root/reference/curdleproofs/curdleproofs/):
//   compute_MSM                       msm_accumulator.py:6-12   -> batched Pippenger (SortDigits,
//                                                                  BucketAccumulate, WindowReduce, Horner)
//   MSMAccumulator.verify's final MSM msm_accumulator.py:60-68  -> same, plus FixedMsm for CRS bases
//   G_L[i] + x*G_R[i] folds           ipa.py:145-146, same_msm.py:124-126 -> FoldPoints
//   G'_i = beta^-(i+1) * G_i          grand_prod.py:66-71       -> MulPoints
//   to/from_compressed_bytes          util.py:27-28,35-36       -> CompressJac / Decompress
//
// Batched Pippenger over B independent MSMs of n terms:
//   window width c, W = ceil(256/c) windows, NB = 2^(c-1) buckets per window (signed digits).
//   0 RecodeDigits      thread = (msm, term): the W signed digits of one scalar
//   1 SortDigits        thread = (msm, window): counting sort of the n signed digits
//   2 BucketAccumulate  thread = (msm, window, bucket): XYZZ mixed adds, registers only
//   3 WindowReduce      thread = (msm, window): running-sum  sum_b (b+1) * S_b
//   4 Horner            thread = msm: c doublings + 1 add per window
// Every body is plain per-thread code, so the identical functors run under the CPU test seam.
#pragma once
#include "g1.cuh"

namespace cpg {

// ---- signed-digit recoding, local form --------------------------------------------------
// k' = k + C with C = sum_{w < W-1} 2^(c-1+cw).  Then for w < W-1: d_w = ((k' >> cw) & (2^c-1)) - 2^(c-1)
// in [-2^(c-1), 2^(c-1)-1], and the top digit d_{W-1} = k' >> c(W-1) in [0, 2^(c-1)] (k < 2^255).
// No carry chain between windows, so each (term, window) pair is independent.
struct Recode {
    uint32_t c, W;
    uint32_t C[8];
};
CPG_HD void recode_add(const Recode& rc, const uint32_t* k, uint32_t* kp) {
    kp[0] = add_cc(k[0], rc.C[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) kp[i] = addc_cc(k[i], rc.C[i]);
    kp[7] = addc(k[7], rc.C[7]);
}
CPG_HD int recode_digit(const Recode& rc, const uint32_t* kp, uint32_t w) {
    uint32_t bit = w * rc.c, limb = bit >> 5, sh = bit & 31;
    uint32_t v = kp[limb] >> sh;
    if (sh + rc.c > 32 && limb < 7) v |= kp[limb + 1] << (32 - sh);
    if (w == rc.W - 1) return (int)v;          // top window: everything that is left, unsigned
    v &= (1u << rc.c) - 1u;
    return (int)v - (int)(1u << (rc.c - 1));
}

struct MsmShape {
    uint32_t B, n, c, W, NB;      // NB = 2^(c-1)
    uint32_t w0, wn;              // windows [w0, w0 + wn) are processed (wn = W: all); per-(msm, window)
                                  // arrays are indexed mw = msm*wn + (w - w0)
    uint64_t base_stride;         // bases of msm m start at m*base_stride (0 = shared by all) ...
    const uint32_t* base_off;     // ... unless this is set: bases of msm m start at base_off[m] (in points)
    uint64_t sc_stride, sc_off;   // scalar of (msm m, term i) = scalars[m*sc_stride + sc_off + i]: a rank of a sharded proof
                                  // evaluates a sub-range of every coefficient row (sc_stride = n, sc_off = 0 otherwise)
};

// 0 --- signed digits of one scalar, all windows ---------------------------------------------
// dig[msm][term][w]: the W digits of a term are contiguous (one thread writes them in a row) and a
// warp of SortDigits threads (consecutive windows of one MSM) reads consecutive int16s: coalesced both ways.
struct RecodeDigits {
    static constexpr const char* kName = "RecodeDigits";
    MsmShape s; Recode rc;
    const uint32_t* scalars;      // [B][n][8] canonical little-endian
    int16_t* dig;                 // [B][n][W] (out)
    CPG_HD void operator()(uint64_t t) const {      // t = msm*n + term
        uint32_t kp[8];
        recode_add(rc, scalars + 8 * ((t / s.n) * s.sc_stride + s.sc_off + t % s.n), kp);
        int16_t* out = dig + t * (uint64_t)s.W;
        for (uint32_t w = 0; w < s.W; w++) out[w] = (int16_t)recode_digit(rc, kp, w);
    }
};

// 1 --- counting sort of the digits of one (msm, window) ----------------------------------
// The per-thread histogram / rank arrays are reached through small views so the same body runs on
// thread-private global arrays (generic launcher, host test seam) or on bank-conflict-free strided
// slices of shared memory (k_sort_digits_smem in cpg_api.cu: element k of thread t at [k*BLOCK + t]).
template <class T> struct LinView { T* p; CPG_HD T& operator[](uint32_t i) const { return p[i]; } };
template <class T> struct StridedView { T* p; uint32_t stride; CPG_HD T& operator[](uint32_t i) const { return p[(size_t)i * stride]; } };

template <class Off, class Rank>
CPG_HD void sort_digits_body(const MsmShape& s, const int16_t* dg, uint32_t* out, Off off, Rank rk, bool want_rank) {
    for (uint32_t b = 0; b <= s.NB; b++) off[b] = 0;
    for (uint32_t i = 0; i < s.n; i++) {           // off[a] = number of terms with |digit| = a
        int d = dg[(uint64_t)i * s.W];
        if (d) off[(uint32_t)(d < 0 ? -d : d)]++;
    }
    uint32_t run = 0;                               // off[a] = END of bucket a-1
    for (uint32_t b = 1; b <= s.NB; b++) { run += off[b]; off[b] = run; }
    const uint32_t total = run;
    for (uint32_t i = s.n; i-- > 0;) {              // fill each bucket from its end (stable)
        int d = dg[(uint64_t)i * s.W];
        if (d) { uint32_t a = (uint32_t)(d < 0 ? -d : d); uint32_t pos = off[a] - 1; off[a] = pos; out[pos] = i | (d < 0 ? 0x80000000u : 0u); }
    }
    // off[a] is now the START of bucket a-1: shift down to off[b] = start of bucket b
    for (uint32_t b = 0; b < s.NB; b++) off[b] = off[b + 1];
    off[s.NB] = total;
    // Rank the buckets by length (insertion sort, NB <= 256).  BucketAccumulate gives one warp the
    // same window and the same rank of 32 different MSMs, whose lengths are tightly concentrated,
    // so its lanes run near-equal trip counts: no global sort, no atomics.
    if (want_rank) {
        for (uint32_t b = 0; b < s.NB; b++) {
            uint32_t len = off[b + 1] - off[b];
            uint32_t j = b;
            while (j > 0) {
                uint32_t pb = rk[j - 1];
                if (off[pb + 1] - off[pb] >= len) break;
                rk[j] = (uint16_t)pb;
                j--;
            }
            rk[j] = (uint16_t)b;
        }
    }
}

struct SortDigits {
    static constexpr const char* kName = "SortDigits";
    MsmShape s;
    const int16_t* dig;           // [B][n][W] from RecodeDigits
    uint32_t* boff;               // [B*W][NB+1] bucket start offsets (out)
    uint32_t* sorted;             // [B*W][n] term index | sign<<31, grouped by bucket (out)
    uint16_t* rank;               // [B*W][NB] buckets of this window ordered by list length, longest first (out; may be null)
    CPG_HD void operator()(uint64_t t) const {
        uint32_t m = (uint32_t)(t / s.wn), w = s.w0 + (uint32_t)(t % s.wn);
        LinView<uint32_t> off{boff + t * (uint64_t)(s.NB + 1)};
        LinView<uint16_t> rk{rank ? rank + t * (uint64_t)s.NB : nullptr};
        sort_digits_body(s, dig + (uint64_t)m * s.n * s.W + w, sorted + t * (uint64_t)s.n, off, rk, rank != nullptr);
    }
};

// 2 --- one bucket: sum of its (signed) bases, mixed XYZZ adds ------------------------------
// Thread -> bucket mapping.  Without `rank`: t = (msm*W + w)*NB + b.  With `rank` (balanced): a warp
// takes ONE window w and ONE rank r of 32 consecutive MSMs (lane = msm): the rank-r buckets of the same
// window of different MSMs have nearly the same length, so the lanes run equal trip counts.  (Mixing
// windows inside a warp is what must be avoided: the top window holds only 255 - c(W-1) bits, i.e. a
// few very long buckets, and one such lane would stall its 31 neighbours.)
struct BucketAccumulate {
    static constexpr const char* kName = "BucketAccumulate";
    MsmShape s;
    const Aff* bases;
    const uint32_t* boff;
    const uint32_t* sorted;
    const uint16_t* rank;         // from SortDigits, or null
    const uint32_t* order;        // from LenScatter (large path: all buckets of the launch by list length), or null
    uint64_t BW;                  // B*W
    Xyzz* buckets;                // [B*W][NB] (out)
    CPG_HD void operator()(uint64_t t) const {
        uint64_t mw; uint32_t b;
        if (order) {
            uint32_t id = order[t];
            mw = id / s.NB; b = id % s.NB;
        } else if (rank) {
            uint32_t lane = (uint32_t)(t % 32), r = (uint32_t)((t / 32) % s.NB);
            uint64_t q = t / (32ull * s.NB);
            uint32_t w = (uint32_t)(q % s.wn);              // window index relative to w0
            uint64_t msm = (q / s.wn) * 32 + lane;
            if (msm >= s.B) return;
            mw = msm * s.wn + w;
            b = rank[mw * s.NB + r];
        } else {
            mw = t / s.NB; b = (uint32_t)(t % s.NB);
        }
        uint32_t m = (uint32_t)(mw / s.wn);
        const uint32_t* off = boff + mw * (uint64_t)(s.NB + 1);
        const uint32_t* lst = sorted + mw * (uint64_t)s.n;
        const Aff* P = bases + (s.base_off ? (uint64_t)s.base_off[m] : (uint64_t)m * s.base_stride);
        Xyzz acc = xyzz_inf();
        uint32_t lo = off[b], hi = off[b + 1];
        if (lo < hi) {
            uint32_t e = lst[lo];
            Aff q = P[e & 0x7fffffffu];
            for (uint32_t j = lo; j < hi; j++) {
                // fetch the next base before the ~10 Fq products of this add (hides the gather latency)
                uint32_t e2 = (j + 1 < hi) ? lst[j + 1] : e;
                Aff q2 = P[e2 & 0x7fffffffu];
                acc = xyzz_add_mixed(acc, cneg(q, (e >> 31) != 0));
                e = e2; q = q2;
            }
        }
        buckets[mw * s.NB + b] = acc;
    }
};

// 2' -- the same bucket sums by BATCHED AFFINE additions ---------------------------------------------
// An affine addition costs one division: lambda = (y2 - y1) / (x2 - x1), x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1.  With Montgomery's trick K independent additions share ONE inversion (field.cuh: ~4 k wide
// multiplies) and each costs 5 products + 1 squaring (forward prefix product, two back-substitution products, lambda,
// lambda^2, y3) instead of the 8 + 2 of the mixed XYZZ addition above.  The additions of a bucket are made independent
// by summing it as a pairwise tree: round r pairs up neighbours inside every bucket, so a thread that owns a run of
// buckets (~500 points: a whole window of a Whisk-size MSM, or a few buckets of a large one) has hundreds of
// independent additions per round, K = BA_K of them per inversion.  Rounds ping-pong between two scratch arrays laid
// out [slot][thread] (a warp's access to one slot is one contiguous run of records); round 0 reads the bases through
// the sorted term list.  Exceptional pairs (P + P, P + (-P), an identity operand) are resolved by flag, they neither
// divide nor stall the others.  A thread whose run would not fit its scratch slots (badly skewed digits) sums its
// buckets with the XYZZ chain instead.  Buckets leave as XYZZ points with ZZ = ZZZ = 1: the reductions are unchanged.
constexpr uint32_t BA_K = 32;        // additions per shared inversion
constexpr uint32_t BA_MIN = 6;       // a round with fewer additions than this is not worth an inversion: finish with XYZZ chains
struct BucketAccumulateAffine {
    static constexpr const char* kName = "BucketAccumulateAffine";
    MsmShape s;
    const Aff* bases;
    const uint32_t* boff;
    const uint32_t* sorted;
    uint64_t BW;                  // (msm, window) pairs of the launch
    uint32_t bpt;                 // buckets per thread (divides NB)
    uint32_t lane_msm;            // 1: bpt = NB and a warp is ONE window of 32 consecutive MSMs (lane = msm; T = ceil32(B) wn), 0: t = (mw, run of buckets)
    uint32_t cap;                 // scratch slots per thread in each of the two arrays
    uint64_t T;                   // threads of the launch (the scratch stride)
    Aff* bufA; Aff* bufB;         // [cap][T]
    Xyzz* buckets;                // [BW][NB] (out)
    struct Src {                  // where round r reads its points
        const Aff* P; const uint32_t* lst; const Aff* buf; uint64_t T, t;
        CPG_HD Aff at(uint32_t i) const {
            if (lst) { const uint32_t e = lst[i]; return cneg(P[e & 0x7fffffffu], (e >> 31) != 0); }
            return buf[(uint64_t)i * T + t];
        }
    };
    // Warp discipline (device): all 32 lanes run the round / batch loops the same number of times (votes decide when a
    // batch is flushed and when a round or the whole tree ends) and meet again after every data-dependent stretch.  Left
    // to themselves the lanes drift apart at the first bucket boundary and never rejoin: ncu showed 9.4 of 32 lanes
    // active per issued instruction and the kernel 4x slower than the XYZZ one.  k_bucket_affine (cpg_api.cu) therefore
    // calls run() for EVERY thread of the grid, `valid` = false for the padding threads, which own no buckets.
#ifdef __CUDA_ARCH__
    static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
    static __device__ __forceinline__ void rejoin() { __syncwarp(); }
#else
    static bool any(bool p) { return p; }
    static void rejoin() {}
#endif
    CPG_HD void operator()(uint64_t t) const { run(t, true); }
    CPG_HD void run(uint64_t t, bool valid) const {
        uint64_t mw = 0; uint32_t b0 = 0;
        if (lane_msm) {
            const uint32_t lane = (uint32_t)(t % 32);
            const uint64_t q = t / 32;
            const uint32_t w = (uint32_t)(q % s.wn);
            const uint64_t msm = (q / s.wn) * 32 + lane;
            if (msm >= s.B) valid = false;
            else mw = msm * s.wn + w;
        } else if (valid) {
            const uint32_t per = s.NB / bpt;
            mw = t / per; b0 = (uint32_t)(t % per) * bpt;
        }
        const uint32_t nbk = valid ? bpt : 0;                                  // buckets of this thread
        const uint32_t m = (uint32_t)(mw / s.wn);
        const uint32_t* off = boff + mw * (uint64_t)(s.NB + 1) + b0;          // off[0 .. bpt]
        const uint32_t* lst = sorted + mw * (uint64_t)s.n;
        const Aff* P = bases + (valid ? (s.base_off ? (uint64_t)s.base_off[m] : (uint64_t)m * s.base_stride) : 0);
        Xyzz* outb = buckets + mw * (uint64_t)s.NB + b0;
        const uint32_t first = valid ? off[0] : 0;
        uint32_t need = 0;
        for (uint32_t b = 0; b < nbk; b++) need += (off[b + 1] - off[b] + 1) >> 1;
        const bool fits = need <= cap;                                         // else: XYZZ chains for this thread's buckets
        Src src{P, lst + first, nullptr, T, t};
        uint32_t r = 0;                                                        // rounds done: bucket b now holds (len_b + 2^r - 1) >> r points
        Aff* out = bufA; Aff* other = bufB;
        for (;; r++) {
            uint32_t adds = 0;
            if (fits) for (uint32_t b = 0; b < nbk; b++) adds += ((off[b + 1] - off[b] + (1u << r) - 1) >> r) >> 1;
            if (!any(adds >= BA_MIN)) break;                                   // the warp goes on while one lane still has a round worth an inversion
            // one round: pairs inside every bucket, BA_K of them per inversion
            Fq pre[BA_K]; uint32_t sa[BA_K], da[BA_K];                          // prefix products; source index of the pair; destination index | flag << 30
            uint32_t cnt = 0, in_pos = 0, out_pos = 0;
            Fq acc = fq_one();
            const uint32_t nb_r = fits ? nbk : 0;
            uint32_t b = 0, j = 0, L = nb_r ? (off[1] - off[0] + (1u << r) - 1) >> r : 0;
            for (;;) {
                // next pair of this lane (or the end of its round)
                bool have = false;
                while (b < nb_r) {
                    if (j + 1 < L) { have = true; break; }
                    if (j < L) out[(uint64_t)(out_pos + (j >> 1)) * T + t] = src.at(in_pos + j);      // odd one out: passes through
                    in_pos += L; out_pos += (L + 1) >> 1;
                    b++; j = 0;
                    if (b < nb_r) L = (off[b + 1] - off[b] + (1u << r) - 1) >> r;
                }
                rejoin();
                const bool anyhave = any(have);
                if (have) {
                    const Aff A = src.at(in_pos + j), B = src.at(in_pos + j + 1);
                    uint32_t flag = 0;                                            // 0 generic, 1 doubling, 2 sum is the identity, 3 an operand is the identity
                    Fq den = fq_one();
                    if (is_inf(A) || is_inf(B)) flag = 3;
                    else {
                        den = sub(B.x, A.x);
                        if (den.is_zero()) {
                            if (A.y == B.y && !A.y.is_zero()) { flag = 1; den = dbl(A.y); }
                            else { flag = 2; den = fq_one(); }
                        }
                    }
                    pre[cnt] = acc;
                    acc = mul(acc, den);
                    sa[cnt] = in_pos + j; da[cnt] = (out_pos + (j >> 1)) | (flag << 30);
                    cnt++; j += 2;
                }
                rejoin();
                if (any(cnt == BA_K) || !anyhave) {                               // lanes still pairing all hold the same count; the others flush what they have
                    Fq inv = fq_inv(acc);
                    for (uint32_t k = BA_K; k-- > 0;) {
                        if (!any(k < cnt)) continue;
                        if (k < cnt) {
                            const Aff A = src.at(sa[k]), B = src.at(sa[k] + 1);
                            const uint32_t flag = da[k] >> 30;
                            Aff R;
                            if (flag == 3) R = is_inf(A) ? B : A;
                            else if (flag == 2) R = aff_inf();
                            else {
                                Fq den, num;
                                if (flag == 0) { den = sub(B.x, A.x); num = sub(B.y, A.y); }
                                else { den = dbl(A.y); Fq xx = sqr(A.x); num = add(dbl(xx), xx); }
                                const Fq li = mul(inv, pre[k]);
                                inv = mul(inv, den);
                                const Fq lam = mul(num, li);
                                R.x = sub(sub(sqr(lam), A.x), B.x);
                                R.y = sub(mul(lam, sub(A.x, R.x)), A.y);
                            }
                            out[(uint64_t)(da[k] & 0x3fffffffu) * T + t] = R;
                        }
                        rejoin();
                    }
                    cnt = 0; acc = fq_one();
                }
                if (!anyhave) break;
            }
            if (fits) { src.lst = nullptr; src.buf = out; }
            Aff* tmp = out; out = other; other = tmp;
        }
        // what is left of every bucket (one point after enough rounds; everything if the run did not fit): XYZZ chain
        uint32_t in_pos = 0;
        const uint32_t rr = fits ? r : 0;
        for (uint32_t b = 0; b < nbk; b++) {
            const uint32_t L = (off[b + 1] - off[b] + (1u << rr) - 1) >> rr;
            Xyzz acc = xyzz_inf();
            for (uint32_t j = 0; j < L; j++) acc = xyzz_add_mixed(acc, src.at(in_pos + j));
            outb[b] = acc;
            in_pos += L;
        }
    }
};

// 3 --- one window: sum_b (b+1) * S_b by running sums ---------------------------------------
struct WindowReduce {
    static constexpr const char* kName = "WindowReduce";
    MsmShape s;
    const Xyzz* buckets;
    Xyzz* wsum;                   // [B*W] (out)
    CPG_HD void operator()(uint64_t t) const {
        const Xyzz* bk = buckets + t * (uint64_t)s.NB;
        Xyzz run = xyzz_inf(), tot = xyzz_inf();
        for (uint32_t b = s.NB; b-- > 0;) {
            run = xyzz_add(run, bk[b]);
            tot = xyzz_add(tot, run);
        }
        wsum[t] = tot;
    }
};

// 4 --- one msm: Horner over its windows ----------------------------------------------------
struct Horner {
    static constexpr const char* kName = "Horner";
    MsmShape s;
    const Xyzz* wsum;
    Jac* out;                     // [B]
    CPG_HD void operator()(uint64_t m) const {
        const Xyzz* ws = wsum + m * (uint64_t)s.W;
        Xyzz acc = ws[s.W - 1];
        for (uint32_t w = s.W - 1; w-- > 0;) {
            for (uint32_t j = 0; j < s.c; j++) acc = xyzz_dbl(acc);
            acc = xyzz_add(acc, ws[w]);
        }
        out[m] = xyzz_to_jac(acc);
    }
};

// ---- large-n path (single big MSMs, n up to 2^20 and beyond) ----------------------------------
// The per-(msm, window) serial counting sort and window reduction above are right for thousands of
// small MSMs; for one MSM of a million terms the same stages are spread over (msm, term) threads
// with atomics and over bucket chunks.  Bucket order inside a list depends on atomic arrival order,
// which is harmless: the group is commutative and results leave as canonical encodings.
#ifdef __CUDA_ARCH__
CPG_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
#else
CPG_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#endif
struct LargeCount {                // thread = (msm, term): cnt[mw][|d|] += 1  (cnt zeroed beforehand)
    static constexpr const char* kName = "LargeCount";
    MsmShape s; const int16_t* dig; uint32_t* cnt;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t m = t / s.n;
        const int16_t* dg = dig + t * (uint64_t)s.W;
        for (uint32_t k = 0; k < s.wn; k++) {
            int d = dg[s.w0 + k];
            if (d) atomic_add_u32(cnt + (m * s.wn + k) * (uint64_t)(s.NB + 1) + (uint32_t)(d < 0 ? -d : d), 1u);
        }
    }
};
// Exclusive scan of the NB counts of every (msm, window) in three short per-thread passes (a window of
// 2^15 buckets scanned by ONE thread was a millisecond of pure latency): chunk totals, a scan of the
// <= NB/SCAN_CH totals, then the starts.  boff[b] = start of bucket b (digit b+1), boff[NB] = list length.
constexpr uint32_t SCAN_CH = 64;
struct LargeScanChunks {           // thread = (mw, chunk)
    static constexpr const char* kName = "LargeScanChunks";
    MsmShape s; uint32_t ch, nch; const uint32_t* cnt; uint32_t* ctot;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t mw = t / nch; uint32_t k = (uint32_t)(t % nch);
        const uint32_t* c = cnt + mw * (uint64_t)(s.NB + 1) + 1 + (uint64_t)k * ch;
        uint32_t sum = 0;
        for (uint32_t i = 0; i < ch; i++) sum += c[i];
        ctot[t] = sum;
    }
};
struct LargeScanTop {              // thread = mw: chunk totals -> exclusive chunk starts
    static constexpr const char* kName = "LargeScanTop";
    uint32_t nch; uint32_t* ctot;
    CPG_HD void operator()(uint64_t mw) const {
        uint32_t* c = ctot + mw * (uint64_t)nch;
        uint32_t run = 0;
        for (uint32_t k = 0; k < nch; k++) { uint32_t v = c[k]; c[k] = run; run += v; }
    }
};
struct LargeScanApply {            // thread = (mw, chunk): bucket starts; the counts become zeroed scatter cursors
    static constexpr const char* kName = "LargeScanApply";
    MsmShape s; uint32_t ch, nch; uint32_t* cnt; const uint32_t* ctot; uint32_t* boff;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t mw = t / nch; uint32_t k = (uint32_t)(t % nch);
        uint32_t* c = cnt + mw * (uint64_t)(s.NB + 1) + 1 + (uint64_t)k * ch;
        uint32_t* o = boff + mw * (uint64_t)(s.NB + 1) + (uint64_t)k * ch;
        uint32_t run = ctot[t];
        for (uint32_t i = 0; i < ch; i++) { o[i] = run; run += c[i]; c[i] = 0; }
        if (k == nch - 1) o[ch] = run;
    }
};
struct LargeScatter {              // thread = (msm, term): claim the next slot of its bucket
    static constexpr const char* kName = "LargeScatter";
    MsmShape s; const int16_t* dig; const uint32_t* boff; uint32_t* cursor; uint32_t* sorted;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t m = t / s.n; uint32_t i = (uint32_t)(t % s.n);
        const int16_t* dg = dig + t * (uint64_t)s.W;
        for (uint32_t k = 0; k < s.wn; k++) {
            int d = dg[s.w0 + k];
            if (!d) continue;
            uint64_t mw = m * s.wn + k;
            uint32_t a = (uint32_t)(d < 0 ? -d : d);
            uint32_t pos = boff[mw * (uint64_t)(s.NB + 1) + a - 1] + atomic_add_u32(cursor + mw * (uint64_t)(s.NB + 1) + a, 1u);
            sorted[mw * (uint64_t)s.n + pos] = i | (d < 0 ? 0x80000000u : 0u);
        }
    }
};
// Buckets ordered by list length, longest first (counting sort over all (msm, window, bucket) of the
// launch): a warp of BucketAccumulate then walks 32 lists of equal length instead of waiting for its
// longest lane (Poisson-distributed lengths cost a third of the kernel otherwise).
constexpr uint32_t LEN_BINS = 1024;
CPG_HD uint32_t len_bin(uint32_t len) { return len < LEN_BINS ? len : LEN_BINS - 1; }
struct LenHist {                   // thread = (mw, bucket)
    static constexpr const char* kName = "LenHist";
    MsmShape s; const uint32_t* boff; uint32_t* hist;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t mw = t / s.NB; uint32_t b = (uint32_t)(t % s.NB);
        const uint32_t* o = boff + mw * (uint64_t)(s.NB + 1) + b;
        atomic_add_u32(hist + len_bin(o[1] - o[0]), 1u);
    }
};
struct LenScan {                   // one thread: histogram -> first slot of every length, longest first
    static constexpr const char* kName = "LenScan";
    uint32_t* hist;
    CPG_HD void operator()(uint64_t) const {
        uint32_t run = 0;
        for (uint32_t l = LEN_BINS; l-- > 0;) { uint32_t v = hist[l]; hist[l] = run; run += v; }
    }
};
struct LenScatter {                // thread = (mw, bucket)
    static constexpr const char* kName = "LenScatter";
    MsmShape s; const uint32_t* boff; uint32_t* hist; uint32_t* order;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t mw = t / s.NB; uint32_t b = (uint32_t)(t % s.NB);
        const uint32_t* o = boff + mw * (uint64_t)(s.NB + 1) + b;
        order[atomic_add_u32(hist + len_bin(o[1] - o[0]), 1u)] = (uint32_t)t;
    }
};
// Window reduction  sum_b (b+1) B_b  in log-depth levels.  With zero-based weights and chunks of CH,
//   R0(X) = sum_i i X_i = sum_c S_c + CH * R0(T),   S_c = sum_j j X[c CH + j],  T_c = sum_j X[c CH + j],
// so every level maps the weighted array to its chunk totals T (the next weighted array) and emits one
// plain-sum array S; earlier levels' S arrays are chunk-summed alongside by their own threads ("roles").
// At length 1:  sum_b (b+1) B_b = T + S^0 + CH_0 (S^1 + CH_1 (S^2 + ...)).  A thread's serial chain is
// 2 CH - 3 additions per level instead of 2 NB for the whole window.
struct ReduceLevel {               // thread = (role, mw, chunk); arrays are [role][BW][len]
    static constexpr const char* kName = "ReduceLevel";
    uint32_t roles, ch, len_in; uint64_t BW; const Xyzz* in; Xyzz* out;
    CPG_HD void operator()(uint64_t t) const {
        uint32_t len_out = len_in / ch;
        uint64_t c = t % len_out, q = t / len_out, mw = q % BW; uint32_t r = (uint32_t)(q / BW);
        const Xyzz* X = in + (r * BW + mw) * (uint64_t)len_in + c * ch;
        if (r == 0) {
            Xyzz run = X[ch - 1], acc = run;
            for (uint32_t j = ch - 1; j-- > 1;) { run = xyzz_add(run, X[j]); acc = xyzz_add(acc, run); }
            out[mw * (uint64_t)len_out + c] = xyzz_add(run, X[0]);
            out[((uint64_t)roles * BW + mw) * len_out + c] = acc;
        } else {
            Xyzz sum = X[0];
            for (uint32_t j = 1; j < ch; j++) sum = xyzz_add(sum, X[j]);
            out[(r * BW + mw) * (uint64_t)len_out + c] = sum;
        }
    }
};
struct ReduceFinal {               // thread = mw; in = [levels + 1][BW][1]; window sums leave as Jacobian
    static constexpr const char* kName = "ReduceFinal";
    uint32_t levels; uint8_t lg_ch[16]; uint64_t BW; const Xyzz* in; Jac* wsum;
    CPG_HD void operator()(uint64_t mw) const {
        Xyzz acc = in[(uint64_t)levels * BW + mw];
        for (uint32_t j = levels - 1; j-- > 0;) {
            for (uint32_t k = 0; k < lg_ch[j]; k++) acc = xyzz_dbl(acc);
            acc = xyzz_add(acc, in[(uint64_t)(j + 1) * BW + mw]);
        }
        wsum[mw] = xyzz_to_jac(xyzz_add(acc, in[mw]));
    }
};
struct XyzzToJac {
    static constexpr const char* kName = "XyzzToJac";
    const Xyzz* in; Jac* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = xyzz_to_jac(in[t]); }
};
struct HornerJac {                 // thread = msm: combine W Jacobian window sums (window-split MSMs)
    static constexpr const char* kName = "HornerJac";
    uint32_t W, c; const Jac* wsum; Jac* out;
    // a SEGMENT of the pass: windows [w_lo, w_hi) from the top down (w_hi = 0 means W); cont != 0 continues from the
    // partial result of the windows above, read from out[m] (the pipelined single MSM, cpg_api.cu)
    uint32_t w_lo = 0, w_hi = 0, cont = 0;
    CPG_HD uint32_t hi() const { return w_hi ? w_hi : W; }
    CPG_HD void operator()(uint64_t m) const {
        const Jac* ws = wsum + m * (uint64_t)W;
        uint32_t w = hi();
        Jac acc;
        if (cont) acc = out[m];
        else acc = ws[--w];
        while (w-- > w_lo) {
            for (uint32_t j = 0; j < c; j++) acc = jac_dbl(acc);
            acc = jac_add(acc, ws[w]);
        }
        out[m] = acc;
    }
};

// ---- fixed-base tables (CRS generators shared by every proof) ------------------------------
// T[i][w][d-1] = d * 2^(c w) * G_i as affine, d = 1..NB.  With them an MSM over the CRS needs
// no doublings and no bucket reduction: n*W mixed adds (SURVEY 8f-2 direction).
struct FixedShape { uint32_t nb, c, W, NB; };

#ifdef __CUDA_ARCH__
CPG_HD int clz32(uint32_t v) { return __clz((int)v); }
#else
CPG_HD int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
#endif
// Table build in two kernels, every inversion shared by Montgomery's trick (round 1 inverted each of the 68.7 M entries
// of the c = 16 table on its own: 1.7 s per table):
//   FixedTableHeads  thread = base i: Q_{i,w} = 2^(c w) G_i for all W windows (255 doublings), made affine with ONE inversion
//   FixedTableSegs   thread = (row (i, w), segment s): entries d = s S + 1 .. (s + 1) S of the row: (s S + 1) Q by
//                    double-and-add, then S - 1 mixed additions of Q, made affine with ONE inversion per segment
struct FixedTableHeads {
    static constexpr const char* kName = "FixedTableHeads";
    FixedShape s;
    const Aff* bases;
    Jac* scratch; Fq* pz;          // [nb][W]
    Aff* heads;                    // [nb][W] (out)
    CPG_HD void operator()(uint64_t i) const {
        Jac* m = scratch + i * s.W; Fq* z = pz + i * s.W; Aff* out = heads + i * s.W;
        Jac acc = to_jac(bases[i]);
        Fq prod = fq_one();
        for (uint32_t w = 0; w < s.W; w++) {
            if (w) for (uint32_t j = 0; j < s.c; j++) acc = jac_dbl(acc);
            m[w] = acc;
            if (!is_inf(acc)) prod = mul(prod, acc.Z);
            z[w] = prod;
        }
        Fq inv = fq_inv(prod);
        for (uint32_t w = s.W; w-- > 0;) {
            Jac q = m[w];
            if (is_inf(q)) { out[w] = aff_inf(); continue; }
            Fq zi = w ? mul(inv, z[w - 1]) : inv;
            inv = mul(inv, q.Z);
            Fq zi2 = sqr(zi);
            Aff r; r.x = mul(q.X, zi2); r.y = mul(q.Y, mul(zi2, zi));
            out[w] = r;
        }
    }
};
struct FixedTableSegs {
    static constexpr const char* kName = "FixedTableSegs";
    FixedShape s;
    uint32_t S, nseg;              // entries per segment, segments per row (S * nseg = NB)
    const Aff* heads;              // [nb*W]
    uint64_t t0;                   // this launch covers threads t0 + u
    Fq* zs; Fq* pz;                // [launch size][S]: Z of every entry / prefix products of the non-zero Z
    Aff* table;                    // [nb*W][NB] (out; holds X, Y of the Jacobian entry until the back-substitution)
    CPG_HD void operator()(uint64_t u) const {
        const uint64_t t = t0 + u, row = t / nseg;
        const uint32_t seg = (uint32_t)(t % nseg);
        const Aff Q = heads[row];
        Aff* out = table + row * (uint64_t)s.NB + (uint64_t)seg * S;
        Fq* z = zs + u * S; Fq* pp = pz + u * S;
        // first entry of the segment: (seg S + 1) Q, bits from the top
        const uint32_t first = seg * S + 1;
        Jac acc = jac_inf();
        for (int b = 31 - clz32(first); b >= 0; b--) {
            acc = jac_dbl(acc);
            if ((first >> b) & 1) acc = jac_add_mixed(acc, Q);
        }
        Fq prod = fq_one();
        for (uint32_t d = 0; d < S; d++) {
            Aff xy; xy.x = acc.X; xy.y = acc.Y;
            const bool inf = is_inf(acc);
            out[d] = inf ? aff_inf() : xy;
            z[d] = inf ? fq_zero() : acc.Z;
            if (!inf) prod = mul(prod, acc.Z);
            pp[d] = prod;
            acc = jac_add_mixed(acc, Q);
        }
        Fq inv = fq_inv(prod);
        for (uint32_t d = S; d-- > 0;) {
            const Fq zd = z[d];
            if (zd.is_zero()) continue;                                  // identity entry: already written as (0, 0)
            Fq zi = d ? mul(inv, pp[d - 1]) : inv;
            inv = mul(inv, zd);
            Fq zi2 = sqr(zi);
            Aff q = out[d];
            Aff r; r.x = mul(q.x, zi2); r.y = mul(q.y, mul(zi2, zi));
            out[d] = r;
        }
    }
};
struct JacToAff {                  // thread = one point (one Fq inversion each)
    static constexpr const char* kName = "JacToAff";
    const Jac* in; Aff* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = jac_to_aff(in[t]); }
};
struct AffToJac {
    static constexpr const char* kName = "AffToJac";
    const Aff* in; Jac* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = to_jac(in[t]); }
};
struct FixedMsmWindow {            // sum_i +-T[i][w][|d|-1] for one (msm, window[, chunk of bases])
    static constexpr const char* kName = "FixedMsmWindow";
    FixedShape s; Recode rc;
    uint32_t B;
    uint32_t nchunk;               // 1, or (few MSMs over many bases: one large proof) the bases split over nchunk threads
    const Aff* table;
    const uint32_t* scalars;       // [B][row_stride][8]; the table's nb bases take entries row_off .. row_off + nb of every row
    uint64_t row_stride, row_off;
    Xyzz* partial;                 // [B*W*nchunk]
    // Thousands of MSMs (nchunk = 1): a warp = one window of 32 consecutive MSMs (lane = msm) - all lanes walk the same
    // bases of the same table segment, and callers that order their MSMs by kind (the prover: output-major) give every
    // lane the same pattern of structurally zero coefficients, so the zero skips do not diverge.
    // A few MSMs over many bases (nchunk > 1: a lone large proof): thread = (msm, window, chunk of <= 64 bases), lanes =
    // consecutive chunks - with lane = msm a warp would hold ONE working lane and the launch is a pure latency chain.
    CPG_HD void operator()(uint64_t t) const {
        uint32_t ch, w; uint64_t m;
        if (nchunk > 1) {
            ch = (uint32_t)(t % nchunk); w = (uint32_t)((t / nchunk) % s.W); m = t / ((uint64_t)nchunk * s.W);
        } else {
            const uint32_t lane = (uint32_t)(t % 32);
            const uint64_t q = t / 32;
            ch = 0; w = (uint32_t)(q % s.W); m = (q / s.W) * 32 + lane;
            if (m >= B) return;
        }
        const uint32_t per = (s.nb + nchunk - 1) / nchunk;
        const uint32_t i0 = ch * per, i1 = i0 + per < s.nb ? i0 + per : s.nb;
        const uint32_t* ks = scalars + (m * row_stride + row_off) * 8;
        Xyzz acc = xyzz_inf();
        uint32_t kp[8];
        for (uint32_t i = i0; i < i1; i++) {
            const uint32_t* k = ks + 8 * (uint64_t)i;
            if ((k[0] | k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7]) == 0) continue;   // zero coefficient: every digit is zero
            recode_add(rc, k, kp);
            int d = recode_digit(rc, kp, w);
            if (!d) continue;
            uint32_t a = (uint32_t)(d < 0 ? -d : d);
            Aff q2 = table[((uint64_t)i * s.W + w) * s.NB + (a - 1)];
            acc = xyzz_add_mixed(acc, cneg(q2, d < 0));
        }
        partial[(m * s.W + w) * nchunk + ch] = acc;
    }
};
struct SumPartialsRagged {         // thread = (msm, group): sum of up to `per` consecutive partials of that MSM's np_in (one level of a reduction tree)
    static constexpr const char* kName = "SumPartials";
    uint32_t np_in, per, np_out; const Xyzz* partial; Xyzz* out;
    CPG_HD void operator()(uint64_t t) const {
        const uint64_t m = t / np_out; const uint32_t g = (uint32_t)(t % np_out);
        const uint32_t lo = g * per, hi = lo + per < np_in ? lo + per : np_in;
        const Xyzz* ps = partial + m * (uint64_t)np_in;
        Xyzz acc = ps[lo];
        for (uint32_t k = lo + 1; k < hi; k++) acc = xyzz_add(acc, ps[k]);
        out[t] = acc;
    }
};
struct CountNonZero {              // thread = scalar: *count += 1 for every non-zero one (work model of the table MSMs, profiling only)
    static constexpr const char* kName = "CountNonZero";
    const uint32_t* scalars; unsigned long long* count;
    uint64_t nb, row_stride, row_off;   // scalar t = (row t / nb, entry row_off + t % nb)
    CPG_HD void operator()(uint64_t t) const {
        const uint32_t* k = scalars + 8 * ((t / nb) * row_stride + row_off + t % nb);
        if ((k[0] | k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7]) != 0) {
#ifdef __CUDA_ARCH__
            atomicAdd(count, 1ULL);
#else
            __atomic_fetch_add(count, 1ULL, __ATOMIC_RELAXED);
#endif
        }
    }
};
struct SumRanks {                  // thread = output t: sum over the ranks' partial sums (after the all-gather of a sharded proof's round)
    static constexpr const char* kName = "SumRanks";
    uint32_t world; uint64_t count, nfirst; const Jac* all; Jac* out_first; Jac* out_second;   // all = [world][count]; t < nfirst -> out_first
    CPG_HD void operator()(uint64_t t) const {
        Jac acc = all[t];
        for (uint32_t r = 1; r < world; r++) acc = jac_add(acc, all[(uint64_t)r * count + t]);
        if (t < nfirst) out_first[t] = acc; else out_second[t - nfirst] = acc;
    }
};
struct SumPartials {               // thread = (msm, group): plain sum of `per` consecutive partials (first stage for long partial lists)
    static constexpr const char* kName = "SumPartials";
    uint32_t per; const Xyzz* partial; Xyzz* out;
    CPG_HD void operator()(uint64_t t) const {
        const Xyzz* ps = partial + t * (uint64_t)per;
        Xyzz acc = ps[0];
        for (uint32_t k = 1; k < per; k++) acc = xyzz_add(acc, ps[k]);
        out[t] = acc;
    }
};
struct SumWindows {                // thread = msm: plain sum of W partials (no doublings)
    static constexpr const char* kName = "SumWindows";
    uint32_t W;
    const Xyzz* partial;
    Jac* out;
    int accumulate;                // 1: out[m] += sum, 0: out[m] = sum
    CPG_HD void operator()(uint64_t m) const {
        const Xyzz* ps = partial + m * (uint64_t)W;
        Xyzz acc = ps[0];
        for (uint32_t w = 1; w < W; w++) acc = xyzz_add(acc, ps[w]);
        Jac r = xyzz_to_jac(acc);
        out[m] = accumulate ? jac_add(out[m], r) : r;
    }
};

// ---- per-base tables for variable bases that enter MANY small MSMs ------------------------------
// The prover's post-shuffle trackers T_i, U_i each appear in 8 of its 30 small MSMs (B_t / B_u and one of
// L or R in every SameMSM round).  With a table of the multiples 1..TS of each base (TS = 2^(c-1), affine),
// such an MSM is W window sums of plain table look-ups (mixed adds, no buckets, no bucket reduction) plus
// one Horner pass: per proof 1984 terms x W x 10 products instead of the bucket method's accumulate +
// 2 NB W full additions per MSM - the reduction was 37 % of the bucket method's work at n = 62.
struct VarTableBuild {             // thread = one base: d * P for d = 1..TS, affine, one inversion per base
    static constexpr const char* kName = "VarTableBuild";
    uint32_t TS; const Aff* bases; uint64_t row_stride, first, count;   // base t = bases[(t / count) * row_stride + first + t % count]
    uint64_t t0;                   // this launch handles bases t0 + u, u < launch size
    Jac* scratch; Fq* pz;          // [launch size][TS] multiples before normalisation / prefix products of their Z
    Aff* table;                    // [nbases][S][TS] (out)
    // S > 1 (a few proofs per call): S SHIFT GROUPS per base, thread = (base, group s): the multiples of 2^(shift_bits s) P.
    // The MSM over such a table folds S windows into one (window s Wp + w' looks up group s), so its Horner pass is
    // Wp = ceil(W / S) windows long instead of W - the dependent doublings that dominate a one-proof call.
    uint32_t S = 1, shift_bits = 0;
    CPG_HD void operator()(uint64_t u) const {
        const uint64_t t = t0 + u, tb = t / S;
        const uint32_t sg = (uint32_t)(t % S);
        Aff P = bases[(tb / count) * row_stride + first + tb % count];
        if (sg) {
            Jac q = to_jac(P);
            for (uint32_t k = 0; k < sg * shift_bits; k++) q = jac_dbl(q);
            P = jac_to_aff(q);
        }
        Jac* m = scratch + u * TS; Fq* z = pz + u * TS; Aff* out = table + t * TS;
        Jac acc = to_jac(P);
        Fq prod = fq_one();
        for (uint32_t d = 0; d < TS; d++) {                 // m[d] = (d + 1) P; infinities (P = O, small-order P) are skipped in the product
            m[d] = acc;
            if (!is_inf(acc)) prod = mul(prod, acc.Z);
            z[d] = prod;
            acc = jac_add_mixed(acc, P);
        }
        Fq inv = fq_inv(prod);                              // prod != 0: a product of non-zero Z
        for (uint32_t d = TS; d-- > 0;) {
            Jac q = m[d];
            if (is_inf(q)) { out[d] = aff_inf(); continue; }
            Fq zi = d ? mul(inv, z[d - 1]) : inv;            // z[d-1] = product of the non-zero Z before d
            inv = mul(inv, q.Z);
            Fq zi2 = sqr(zi);
            Aff r; r.x = mul(q.X, zi2); r.y = mul(q.Y, mul(zi2, zi));
            out[d] = r;
        }
    }
};
struct VarTableMsmWindow {         // sum_i +-tab[base i of msm m][|d_w| - 1] for one (msm, window); lane = msm as in FixedMsmWindow
    static constexpr const char* kName = "VarTableMsmWindow";
    uint32_t nb, TS, W; Recode rc; uint32_t B;
    const Aff* table; const uint32_t* tab_off;   // tab_off[m]: first base (in bases, not entries) of msm m's nb consecutive tables
    const uint32_t* scalars;       // [B][nb][8]
    Xyzz* partial;                 // [B*W*nchunk]
    uint32_t nchunk = 1;           // > 1 (a few MSMs: one proof per call): thread = (msm, window, chunk of the bases), as FixedMsmWindow
    uint32_t S = 1, Wp = 0;        // shift groups of the table (VarTableBuild) and windows per group; Wp = 0 means W (no groups)
    CPG_HD void operator()(uint64_t t) const {
        const uint32_t WP = Wp ? Wp : W;
        uint32_t w, ch = 0; uint64_t m;
        if (nchunk > 1) {
            ch = (uint32_t)(t % nchunk); w = (uint32_t)((t / nchunk) % WP); m = t / ((uint64_t)nchunk * WP);
        } else {
            const uint32_t lane = (uint32_t)(t % 32);
            w = (uint32_t)((t / 32) % WP); m = (t / (32ull * WP)) * 32 + lane;
        }
        if (m >= B) return;
        const uint32_t* ks = scalars + m * nb * 8;
        const Aff* tab = table + (uint64_t)tab_off[m] * S * TS;
        Xyzz acc = xyzz_inf();
        uint32_t kp[8];
        const uint32_t per = (nb + nchunk - 1) / nchunk;
        const uint32_t i0 = ch * per, i1 = i0 + per < nb ? i0 + per : nb;
        for (uint32_t i = i0; i < i1; i++) {
            const uint32_t* k = ks + 8 * (uint64_t)i;
            if ((k[0] | k[1] | k[2] | k[3] | k[4] | k[5] | k[6] | k[7]) == 0) continue;
            recode_add(rc, k, kp);
            for (uint32_t sg = 0; sg < S; sg++) {
                const uint32_t wf = sg * WP + w;                 // the full-width window this group's entry stands for
                if (wf >= W) break;
                int d = recode_digit(rc, kp, wf);
                if (!d) continue;
                uint32_t a = (uint32_t)(d < 0 ? -d : d);
                acc = xyzz_add_mixed(acc, cneg(tab[((uint64_t)i * S + sg) * TS + (a - 1)], d < 0));
            }
        }
        partial[(m * WP + w) * nchunk + ch] = acc;
    }
};

// ---- element-wise kernels (vector scalar-mul, fold, group law, serialisation) ---------------
// Two kernels, not a runtime flag: the subgroup check is a scalar multiplication, and merely carrying its code makes the
// kernel need 168+ registers; the unchecked decoder (every tracker and proof point, cp/util.py:35-36) fits 96.
struct Decompress {
    static constexpr const char* kName = "Decompress";
    const uint8_t* in; Aff* out; uint8_t* err;
    CPG_HD void operator()(uint64_t t) const {
        Aff a;
        int e = aff_decompress(in + 48 * t, false, &a);
        out[t] = a;
        err[t] = (uint8_t)e;
    }
};
struct DecompressChecked {
    static constexpr const char* kName = "Decompress";
    const uint8_t* in; Aff* out; uint8_t* err;
    CPG_HD void operator()(uint64_t t) const {
        Aff a;
        int e = aff_decompress(in + 48 * t, true, &a);
        out[t] = a;
        err[t] = (uint8_t)e;
    }
};
struct CompressJac {
    static constexpr const char* kName = "CompressJac";
    const Jac* in; uint8_t* out;
    CPG_HD void operator()(uint64_t t) const { aff_compress(jac_to_aff(in[t]), out + 48 * t); }
};
struct CompressAff {
    static constexpr const char* kName = "CompressAff";
    const Aff* in; uint8_t* out;
    CPG_HD void operator()(uint64_t t) const { aff_compress(in[t], out + 48 * t); }
};
struct AddPoints {                 // op 0: a+b, 1: a-b
    static constexpr const char* kName = "AddPoints";
    const Jac* a; const Jac* b; Jac* out; int op;
    CPG_HD void operator()(uint64_t t) const { out[t] = jac_add(a[t], op ? neg(b[t]) : b[t]); }
};
struct NegPoints {
    static constexpr const char* kName = "NegPoints";
    const Jac* a; Jac* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = neg(a[t]); }
};
struct EqPoints {
    static constexpr const char* kName = "EqPoints";
    const Jac* a; const Jac* b; uint8_t* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = jac_eq(a[t], b[t]) ? 1 : 0; }
};
struct IsInfPoints {
    static constexpr const char* kName = "IsInfPoints";
    const Jac* a; uint8_t* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = is_inf(a[t]) ? 1 : 0; }
};
struct MulPoints {                 // out[t] = k[t / group] * p[t]   (group = 1: one scalar per point)
    static constexpr const char* kName = "MulPoints";
    const Jac* p; const uint32_t* k; uint64_t group; Jac* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = jac_mul(p[t], k + 8 * (t / group)); }
};
struct FoldPoints {                // out[r][i] = L[r][i] + x[r] * R[r][i],  r < rows, i < m
    static constexpr const char* kName = "FoldPoints";
    const Jac* L; const Jac* R; const uint32_t* x; uint64_t m; Jac* out;
    CPG_HD void operator()(uint64_t t) const { out[t] = jac_add(L[t], jac_mul(R[t], x + 8 * (t / m))); }
};

// ---- device-resident cache of decompressed points, keyed by their 48-byte encodings ----------------------------
// In Whisk the pre-shuffle trackers of one shuffle proof are post-shuffle trackers of an earlier one
// (cp/whisk_interface.py:96-100 decodes all 4 ell of them per call): a verifier that has seen a tracker already knows its
// square root.  Open-addressing table, linear probing, 64-bit fingerprints backed by a full key comparison; it also
// removes duplicates INSIDE a batch (the first thread to claim a slot decompresses, the others copy).
//   CacheClaim       thread = cached point: hash, probe, claim an empty slot (OWN) or find the key's slot (FOLLOW)
//   CacheResolve     thread = point: FOLLOW -> compare the key, settle HIT / wait-for-owner / decompress-uncached;
//                    every point that must be decompressed is appended to a compact work list (warp-aggregated atomics)
//   DecompressList   thread = work-list entry: square root; an owner also fills its slot and publishes it
//   CacheFill        thread = cached point: hits and followers copy the slot's value
// Several sub-batches run these on their own streams against the one table.  A slot is trusted only when `ready` is set
// (published after the value) or when it was claimed by THIS launch sequence (owner == launch id: the owner's
// DecompressList precedes our CacheFill in stream order); anything else - a torn key, a slot another stream is still
// filling - falls back to decompressing the point uncached.  Verdicts never depend on who wins a race.
#ifdef __CUDA_ARCH__
CPG_HD unsigned long long atomic_cas_u64(unsigned long long* p, unsigned long long expect, unsigned long long v) { return atomicCAS(p, expect, v); }
CPG_HD unsigned long long load_u64_volatile(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
CPG_HD uint32_t load_u32_volatile(const uint32_t* p) { return *(const volatile uint32_t*)p; }
CPG_HD void publish_u32(uint32_t* p, uint32_t v) { __threadfence(); *(volatile uint32_t*)p = v; }
// index = counter++ for every thread with pred set, ONE atomic per warp
CPG_HD uint32_t warp_agg_inc(uint32_t* counter, bool pred) {
    const unsigned active = __activemask();
    const unsigned m = __ballot_sync(active, pred);
    if (!pred) return 0;
    const int lane = (int)(threadIdx.x & 31), leader = __ffs((int)m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
#else
CPG_HD unsigned long long atomic_cas_u64(unsigned long long* p, unsigned long long expect, unsigned long long v) {
    __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return expect;
}
CPG_HD unsigned long long load_u64_volatile(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
CPG_HD uint32_t load_u32_volatile(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
CPG_HD void publish_u32(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
CPG_HD uint32_t warp_agg_inc(uint32_t* counter, bool pred) { return pred ? __atomic_fetch_add(counter, 1u, __ATOMIC_RELAXED) : 0; }
#endif

struct PointCache {
    uint32_t lg, max_probe;            // 2^lg slots
    unsigned long long* tag;           // [slots] 0 = empty, else fingerprint | 1
    unsigned long long* owner;         // [slots] launch id of the claimer
    uint32_t* ready;                   // [slots] 1 once val / verr are valid
    uint32_t* key;                     // [slots][12] the 48 encoded bytes
    Aff* val;                          // [slots]
    uint8_t* verr;                     // [slots] decode error code of the encoding (invalid encodings are cached too)
    uint32_t* stats;                   // [0] slots claimed  [1] lookups  [2] lookups served from the table  (since the last reset)
};
enum { PC_NONE = 0, PC_OWN = 1, PC_FOLLOW = 2, PC_COPY = 3, PC_UNCACHED = 4 };

struct PointSel {                      // which points of a [rows][row_pts] array of 48-byte encodings take part: the first `cached` of every row
    uint64_t row_pts; uint32_t cached;
    CPG_HD uint64_t index(uint64_t t) const { return (t / cached) * row_pts + t % cached; }
};
CPG_HD unsigned long long pc_hash(const uint32_t* w) {
    unsigned long long h = 0x9e3779b97f4a7c15ULL;
    for (int i = 0; i < 12; i += 2) {
        unsigned long long v = (unsigned long long)w[i] | ((unsigned long long)w[i + 1] << 32);
        h = (h ^ v) * 0xff51afd7ed558ccdULL;
        h ^= h >> 32;
    }
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 29;
    return h;
}
struct CacheClaim {
    static constexpr const char* kName = "CacheClaim";
    PointCache pc; PointSel sel; unsigned long long launch_id;
    const uint32_t* in;                // encodings, 12 words each (16-byte aligned rows)
    uint8_t* role; uint32_t* ref;      // per point (indexed like `in`)
    CPG_HD void operator()(uint64_t t) const {
        const uint64_t i = sel.index(t);
        uint32_t w[12];
        for (int k = 0; k < 12; k++) w[k] = in[12 * i + k];
        const unsigned long long h = pc_hash(w), tg = h | 1ULL;
        const uint32_t mask = (1u << pc.lg) - 1u;
        uint32_t slot = (uint32_t)(h >> 20) & mask;
        uint8_t r = PC_UNCACHED; uint32_t at = 0;
        for (uint32_t p = 0; p < pc.max_probe; p++, slot = (slot + 1) & mask) {
            unsigned long long cur = load_u64_volatile(pc.tag + slot);
            if (cur == 0) {
                cur = atomic_cas_u64(pc.tag + slot, 0ULL, tg);
                if (cur == 0) {
                    pc.owner[slot] = launch_id;
                    for (int k = 0; k < 12; k++) pc.key[12 * (uint64_t)slot + k] = w[k];
                    r = PC_OWN; at = slot;
                    break;
                }
            }
            if (cur == tg) { r = PC_FOLLOW; at = slot; break; }
        }
        role[i] = r; ref[i] = at;
    }
};
struct CacheResolve {
    static constexpr const char* kName = "CacheResolve";
    PointCache pc; PointSel sel; unsigned long long launch_id;
    uint32_t use_cache;                // 0: every point goes to the work list
    uint32_t row_take;                 // points of every row that are decoded at all (the rest of the row is left alone)
    const uint32_t* in; uint8_t* role; const uint32_t* ref;
    uint32_t* todo; uint32_t* count;   // work list (out) and its length
    CPG_HD void operator()(uint64_t t) const {
        const uint64_t row = t / row_take, j = t % row_take, i = row * sel.row_pts + j;
        bool need = true;
        uint32_t claimed = 0, served = 0;
        if (use_cache && j < sel.cached) {
            uint8_t r = role[i];
            if (r == PC_FOLLOW) {
                const uint32_t slot = ref[i];
                bool same = true;
                for (int k = 0; k < 12; k++) same = same && pc.key[12 * (uint64_t)slot + k] == in[12 * i + k];
                if (same && (load_u32_volatile(pc.ready + slot) != 0 || pc.owner[slot] == launch_id)) r = PC_COPY;
                else r = PC_UNCACHED;                                    // fingerprint collision, or a slot another stream is still filling
                role[i] = r;
            }
            need = r != PC_COPY;
            claimed = r == PC_OWN; served = r == PC_COPY;
        } else if (j < sel.cached) role[i] = PC_NONE;
        const uint32_t at = warp_agg_inc(count, need);
        if (need) todo[at] = (uint32_t)i;
        if (use_cache) {                                                 // statistics (three warp-aggregated counters)
            warp_agg_inc(pc.stats + 0, claimed != 0);
            warp_agg_inc(pc.stats + 1, j < sel.cached);
            warp_agg_inc(pc.stats + 2, served != 0);
        }
    }
};
struct DecompressList {                // launched over an upper bound of the list length; threads beyond *count leave at once
    static constexpr const char* kName = "Decompress";                   // same work as Decompress: reported under its name
    PointCache pc; uint32_t use_cache; uint32_t cached_per_row; uint64_t row_pts;
    const uint8_t* in; const uint32_t* todo; const uint32_t* count;
    const uint8_t* role; const uint32_t* ref;
    Aff* out; uint8_t* err;
    CPG_HD void operator()(uint64_t t) const {
        if (t >= *count) return;
        const uint64_t i = todo[t];
        Aff a;
        const int e = aff_decompress(in + 48 * i, false, &a);
        out[i] = a; err[i] = (uint8_t)e;
        if (use_cache && (i % row_pts) < cached_per_row && role[i] == PC_OWN) {
            const uint32_t slot = ref[i];
            pc.val[slot] = a; pc.verr[slot] = (uint8_t)e;
            publish_u32(pc.ready + slot, 1u);
        }
    }
};
struct CacheFill {
    static constexpr const char* kName = "CacheFill";
    PointCache pc; PointSel sel; const uint8_t* role; const uint32_t* ref; Aff* out; uint8_t* err;
    CPG_HD void operator()(uint64_t t) const {
        const uint64_t i = sel.index(t);
        if (role[i] != PC_COPY) return;
        const uint32_t slot = ref[i];
        out[i] = pc.val[slot]; err[i] = pc.verr[slot];
    }
};

// ---- Fr vector ops on canonical little-endian words (K7) ------------------------------------
struct FrBinary {                  // op 0 add, 1 sub, 2 mul
    static constexpr const char* kName = "FrBinary";
    const uint32_t* a; const uint32_t* b; uint32_t* out; int op;
    CPG_HD void operator()(uint64_t t) const {
        Fr x, y, z;
        for (int i = 0; i < 8; i++) { x.l[i] = a[8 * t + i]; y.l[i] = b[8 * t + i]; }
        if (op == 0) z = add(x, y);
        else if (op == 1) z = sub(x, y);
        else z = from_mont(mul(to_mont(x), to_mont(y)));
        for (int i = 0; i < 8; i++) out[8 * t + i] = z.l[i];
    }
};
struct FrInverse {                 // 0 -> 0 (cp/util.py:51-54 relies on a value coming back)
    static constexpr const char* kName = "FrInverse";
    const uint32_t* a; uint32_t* out;
    CPG_HD void operator()(uint64_t t) const {
        Fr x;
        for (int i = 0; i < 8; i++) x.l[i] = a[8 * t + i];
        Fr z = x.is_zero() ? x : from_mont(fr_inv(to_mont(x)));
        for (int i = 0; i < 8; i++) out[8 * t + i] = z.l[i];
    }
};

}  // namespace cpg

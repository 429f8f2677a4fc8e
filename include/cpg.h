/* cpg.h - C ABI of the B200-native curdleproofs G1 hot path (libcpg.so).
 *
 * This is the drop-in boundary for the arithmetic the reference obtains from the external
 * wheel py_arkworks_bls12381 0.3.5 (Rust/PyO3).  The reference binds it as Python classes
 *   G1Point  /root/reference/curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:5-30
 *   Scalar   /root/reference/curdleproofs/py_arkworks_bls12381-stubs/__init__.pyi:32-54
 * and calls it one group element at a time.  A maintainer binds THIS header with ctypes/cffi
 * (INTEGRATION.md shows the stub); every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - one process per GPU; cpg_init(device) selects it.  All work is enqueued on the library's
 *     current stream (cpg_set_stream) and is asynchronous unless stated; cpg_sync() waits.
 *   - "d_" arguments are DEVICE pointers obtained from cpg_malloc.  Host buffers cross only
 *     through cpg_h2d / cpg_d2h.
 *   - device formats (little-endian u32 limbs, Fq in Montgomery form R = 2^384):
 *       affine   point: 96 B  = x | y          (x = y = 0 encodes the identity)
 *       jacobian point: 144 B = X | Y | Z      (Z = 0 encodes the identity)
 *       scalar        : 32 B canonical little-endian integer < r (NOT Montgomery), as on the wire
 *       compressed    : 48 B ZCash/IETF encoding, as on the wire (cp/util.py:27-28)
 *   - every function returns 0 on success; non-zero is an error whose text cpg_last_error()
 *     returns.  Encoding errors are reported per element in the err arrays (0 = ok) so that a
 *     bad lane never aborts a batch; the Python surface turns them into ValueError.
 *   - there is no CPU fallback: without a CUDA device cpg_init fails and nothing else works.
 */
#ifndef CPG_H
#define CPG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPG_AFF_BYTES 96
#define CPG_JAC_BYTES 144
#define CPG_SCALAR_BYTES 32
#define CPG_COMPRESSED_BYTES 48

/* ---- runtime ------------------------------------------------------------------------------ */
int cpg_init(int device);                       /* idempotent */
int cpg_device_count(void);
const char* cpg_last_error(void);
const char* cpg_backend(void);                  /* "cuda-sm_100a" for the product library */
int cpg_set_stream(void* cuda_stream);          /* cudaStream_t, NULL = library default stream */
int cpg_sync(void);
void* cpg_malloc(size_t bytes);                 /* NULL on failure */
int cpg_free(void* d_ptr);
int cpg_memset(void* d_ptr, int value, size_t bytes);
int cpg_h2d(void* d_dst, const void* h_src, size_t bytes);   /* async on the current stream */
int cpg_d2h(void* h_dst, const void* d_src, size_t bytes);   /* synchronises before returning */
int cpg_d2d(void* d_dst, const void* d_src, size_t bytes);
void* cpg_host_alloc(size_t bytes);             /* pinned host memory for the e2e path */
int cpg_host_free(void* h_ptr);
/* device timers on the current stream (CUDA events) */
int cpg_timer_start(void);
int cpg_timer_stop(float* ms);                  /* synchronises */
uint64_t cpg_launch_count(void);                /* kernels launched by this library so far */
/* per-kernel device time: when enabled every launch is bracketed by a CUDA event pair on its
 * stream; report() synchronises and writes JSON {"Kernel": {"ms":..,"launches":..,"threads":..}, ...,
 * "_counters": {"fixed_msm_terms": n, "var_table_msm_terms": n}} - the counters are the non-zero coefficients the
 * table-lookup MSMs met since the last reset (each costs one mixed addition per window: their work model) */
int cpg_profile_enable(int on);
int cpg_profile_reset(void);
int cpg_profile_report(char* buf, size_t cap);

/* ---- serialisation -------------------------------------------------------------------------
 * replaces G1Point.from_compressed_bytes (stub :19, check_subgroup=1) and
 * from_compressed_bytes_unchecked (stub :22; cp/util.py:35-36, cp/msm_accumulator.py:65) */
int cpg_g1_decompress(const uint8_t* d_in48, size_t k, int check_subgroup, void* d_out_aff, uint8_t* d_err);
/* replaces G1Point.to_compressed_bytes (stub :30; cp/util.py:27-28) */
int cpg_g1_compress(const void* d_jac, size_t k, uint8_t* d_out48);
int cpg_g1_compress_aff(const void* d_aff, size_t k, uint8_t* d_out48);
int cpg_g1_aff_to_jac(const void* d_aff, size_t k, void* d_out_jac);
int cpg_g1_jac_to_aff(const void* d_jac, size_t k, void* d_out_aff);
int cpg_g1_generator(void* d_out_jac);          /* G1Point()            stub :6  */
int cpg_g1_identity(void* d_out_jac);           /* G1Point.identity()   stub :25 */

/* ---- element-wise group law over k points ---------------------------------------------------
 * replaces G1Point.__add__/__sub__/__neg__/__eq__ (stub :7-16) */
int cpg_g1_add(const void* d_a_jac, const void* d_b_jac, size_t k, void* d_out_jac);
int cpg_g1_sub(const void* d_a_jac, const void* d_b_jac, size_t k, void* d_out_jac);
int cpg_g1_neg(const void* d_a_jac, size_t k, void* d_out_jac);
int cpg_g1_eq(const void* d_a_jac, const void* d_b_jac, size_t k, uint8_t* d_out);
int cpg_g1_is_identity(const void* d_a_jac, size_t k, uint8_t* d_out);
/* replaces G1Point.__mul__(Scalar) (stub :10); out[i] = scalars[i / group] * p[i]
 * (group = 1: one scalar per point; group = m: one scalar per row of m points, the
 * vector scalar-mul  G'_i = beta^-(i+1) G_i  of cp/grand_prod.py:66-71 uses group = 1) */
int cpg_g1_mul(const void* d_p_jac, const uint8_t* d_scalars, size_t k, size_t group, void* d_out_jac);
/* IPA / SameMSM generator folding  out[r][i] = L[r][i] + x[r] * R[r][i]
 * (cp/ipa.py:145-146, cp/same_msm.py:124-126), rows x m points */
int cpg_g1_fold(const void* d_L_jac, const void* d_R_jac, const uint8_t* d_x, size_t rows, size_t m, void* d_out_jac);

/* ---- multi-scalar multiplication -------------------------------------------------------------
 * replaces compute_MSM (cp/msm_accumulator.py:6-12) and G1Point.multiexp_unchecked (stub :28):
 * B independent MSMs of n terms each in one call.  bases of MSM b start at
 * d_bases_aff + b*base_stride points (base_stride = 0: all MSMs share one base vector);
 * scalars are [B][n] x 32 B.  window = 0 picks the window width from n. */
int cpg_g1_msm_batched(const void* d_bases_aff, size_t base_stride, const uint8_t* d_scalars,
                       size_t B, size_t n, int window, void* d_out_jac);
/* same, with an explicit start (in points) of every MSM's base vector: bases of MSM b are
 * d_bases_aff[d_base_off[b] ...] - lets MSMs over different sub-vectors share one call */
int cpg_g1_msm_batched_off(const void* d_bases_aff, const uint32_t* d_base_off, const uint8_t* d_scalars,
                           size_t B, size_t n, int window, void* d_out_jac);
/* Which kernel pipeline cpg_g1_msm_batched uses: 0 (default) by shape - thousands of small MSMs run one
 * thread per (msm, window) for the sort/reduce stages, few or large MSMs run one thread per term with
 * atomics, length-ordered buckets and a level-wise window reduction; 1 / 2 force either (tests, tuning). */
int cpg_msm_force_path(int path);
/* How the bucket accumulation adds: 0 (default) = one mixed XYZZ addition per term (8 products + 2 squarings), one thread
 * per bucket; 1 = batched affine additions - a thread sums a run of buckets as pairwise trees, 32 independent additions
 * sharing one Fq inversion (5 products + 1 squaring per addition: 39 % fewer multiply-accumulates).  Results are
 * identical.  Measured on B200 the affine kernel is SLOWER (it issues more instructions per addition and stalls on its
 * gathers: DESIGN.md section 4, profiles/r02_ncu_bucket_affine_*), so it is kept as a tested alternative, not the default. */
int cpg_msm_set_accumulate(int mode);
/* Window split of ONE large MSM (BASELINE configs 2 and 5): the W = cpg_msm_window_count
 * windows are independent; a rank computes the Jacobian window sums S_w of its slice, the slices are
 * exchanged by one small all-gather (W * 144 B in total), and every rank finishes with
 * sum_w 2^(c w) S_w.  `window` must be the same on every rank (cpg_msm_pick_window(n)). */
int cpg_msm_pick_window(size_t n);
int cpg_msm_pick_window_batched(size_t B, size_t n);   /* what cpg_g1_msm_batched(window = 0) uses for B MSMs of n terms */
int cpg_msm_window_count(size_t n, int window);
int cpg_g1_msm_window_sums(const void* d_bases_aff, const uint8_t* d_scalars, size_t n, int window,
                           int w_begin, int w_end, void* d_out_jac);
int cpg_g1_msm_combine_windows(const void* d_wsums_jac, int window, void* d_out_jac);
/* ---- multi-GPU: the communicator and the path's one collective, inside the library ------------------
 * One process per GPU.  Rank 0 obtains an id (cpg_comm_unique_id, 128 bytes), the host plumbing hands it to the other
 * ranks (any channel: curdleproofs_pie_b200/comm.py uses a TCP socket on MASTER_ADDR), every rank calls
 * cpg_comm_init.  NCCL (libnccl.so.2, dlopen'ed) then moves DEVICE buffers over NVLink/NVSwitch on the library's
 * current stream - no torch, no host bounce.  Without cpg_comm_init the world is 1 and every call degenerates to
 * its single-GPU form.  Replaces nothing in the reference (it is single-process); SURVEY 8e specifies it. */
int cpg_comm_unique_id(uint8_t* out128);
int cpg_comm_init(int rank, int world, const uint8_t* id128);
int cpg_comm_free(void);
int cpg_comm_rank(void);
int cpg_comm_world(void);
int cpg_comm_nccl_version(void);                 /* e.g. 22809; 0 if NCCL cannot be loaded */
/* all-gather of bytes_per_rank bytes from every rank into d_recv[world * bytes_per_rank] (device buffers) */
int cpg_comm_allgather(const void* d_send, void* d_recv, size_t bytes_per_rank);
/* ONE n-term MSM (multiexp_unchecked, stub :28) over the communicator: this rank's slice of the Pippenger windows,
 * ONE ncclAllGather of the window sums (ceil(W/world) x 144 B per rank), Horner on every rank; all ranks hold all
 * bases and scalars and get the same Jacobian result.  window = 0: cpg_msm_pick_window(n). */
int cpg_g1_msm_sharded(const void* d_bases_aff, const uint8_t* d_scalars, size_t n, int window, void* d_out_jac);

/* fixed-base tables for generators shared by every proof (the CRS, cp/crs.py:19-36):
 * T[i][w][d] = (d+1) 2^(c w) G_i.  Returns NULL on failure. */
void* cpg_fixed_table_create(const void* d_bases_aff, size_t nb, int window);
int cpg_fixed_table_free(void* table);
size_t cpg_fixed_table_bytes(const void* table);
/* out[b] (+)= sum_i scalars[b][i] * G_i ; accumulate = 1 adds into d_out_jac */
int cpg_g1_msm_fixed_batched(const void* table, const uint8_t* d_scalars, size_t B, int accumulate, void* d_out_jac);

/* ---- Fr vector ops (Scalar.__add__/__sub__/__mul__/inverse, stub :32-54) on k canonical scalars */
int cpg_fr_add(const uint8_t* d_a, const uint8_t* d_b, size_t k, uint8_t* d_out);
int cpg_fr_sub(const uint8_t* d_a, const uint8_t* d_b, size_t k, uint8_t* d_out);
int cpg_fr_mul(const uint8_t* d_a, const uint8_t* d_b, size_t k, uint8_t* d_out);
int cpg_fr_inverse(const uint8_t* d_a, size_t k, uint8_t* d_out);   /* inverse(0) = 0, cp/util.py:51-54 */

/* ---- Fiat-Shamir transcript ---------------------------------------------------------------------
 * The library's own STROBE-128 / Merlin implementation (the one its prover and verifier run per proof), driven by a
 * byte script; replaces merlin_transcripts.MerlinTranscript (merlin_transcripts/merlin_transcript.py:6-24) over
 * Strobe128 (strobe.py:16-107).  Record: op u8 | more u8 | label_len u16 LE | n u32 LE | label | data[n]
 * (the two output ops carry no data; n = bytes wanted):
 *   0 STROBE init(data = protocol label)   1 meta_ad(data, more)   2 ad(data, more)   3 prf(n, more) -> n bytes
 *   4 key(data, more)   5 MerlinTranscript(data = label)   6 append_message(label, data)   7 challenge_bytes(label, n)
 * Outputs are concatenated into `out`.  on_device = 0 runs the script on the calling thread (host-only; needs no
 * cpg_init), 1 in a one-thread kernel, 2 on one warp in lock-step with the warp-cooperative Keccak permutation (the two
 * placements the batched verifier uses per proof). */
int cpg_merlin_script(const uint8_t* script, size_t len, int on_device, uint8_t* out, size_t out_cap, size_t* out_len);

/* stateful host-side form: MerlinTranscript(label) / append_message / challenge_bytes (merlin_transcript.py:6-24);
 * dropin/merlin_transcripts binds these so that the unmodified reference's Fiat-Shamir runs off Python */
void* cpg_merlin_new(const uint8_t* label, size_t n);
void* cpg_merlin_clone(const void* transcript);
int cpg_merlin_free(void* transcript);
int cpg_merlin_append(void* transcript, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t n);
int cpg_merlin_challenge(void* transcript, const uint8_t* label, size_t label_len, uint8_t* out, size_t n);

/* ---- batched shuffle-proof verification -------------------------------------------------------
 * Replaces, for B proofs at once, IsValidWhiskShuffleProof (cp/whisk_interface.py:74-108) ->
 * CurdleProofsProof.verify (cp/curdleproofs.py:162-248) and everything below it.  Every group
 * operation runs on the GPU (decompress, D / A', one MSM per proof); the transcript and the Fr
 * coefficient algebra run per proof either on the GPU or on `host_threads` host threads
 * (cpg_verifier_set_transcript; 0 threads = all cores).
 *   crs_bytes : (ell + n_blinders + 5) * 48 B = CurdleproofsCrs.to_bytes (cp/crs.py:93-102)
 *   inputs    : [B][4*ell*48]  vec_R | vec_S | vec_T | vec_U  (tracker halves, whisk_interface.py:96-100)
 *   proofs    : [B][cpg_verifier_proof_bytes]  M | proof      (WhiskShuffleProof.to_bytes, :57-61)
 *   verdicts  : [B], 1 = the reference would return True */
void* cpg_verifier_create(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads);
/* BASELINE config 5's verify side: every proof over ALL ranks of the communicator (cpg_comm_init).  Every rank calls
 * this and every later cpg_verify_batch with identical arguments and gets identical verdicts: each proof's MSM terms
 * (CRS bases through a table of this rank's block only, trackers / proof points through the bucket method) are split
 * over the ranks and the 2 partial sums per proof cross by ONE all-gather inside the library. */
void* cpg_verifier_create_sharded(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads);
int cpg_verifier_free(void* verifier);
size_t cpg_verifier_proof_bytes(const void* verifier);
size_t cpg_verifier_input_bytes(const void* verifier);
int cpg_verifier_set_window(void* verifier, int var_window);
/* where the Fiat-Shamir transcript + coefficient algebra run: 0 = on `host_threads` host threads (the reference's
 * placement: right when every thread gets at most one proof - a CPU core runs the sequential Keccak chain ~20x faster
 * than one GPU thread); 1 = one proof per GPU THREAD (only wire bytes cross PCIe; fewest issue slots, but a latency of
 * ~31 ms at n = 128 whatever the batch size: right when other sub-batches' kernels hide it, i.e. thousands of proofs);
 * 3 = one proof per GPU WARP (the 32 lanes share every Keccak permutation through shuffles; 32x the issue slots: right
 * for tens to ~1000 proofs); 2 (default) = by batch size: host while every host thread gets at most one proof, warp per
 * proof up to 1024 proofs, thread per proof beyond */
int cpg_verifier_set_transcript(void* verifier, int mode);
/* at most this many sub-batches in flight, each on its own set of CUDA streams (1..8, default 8; device transcript only;
 * a sub-batch holds at least 1024 proofs): the host stages the wire bytes of sub-batch k + 1 while the GPU works on
 * sub-batch k, the transcript kernels run on high-priority side streams, and every sub-batch's decompression is enqueued
 * before any MSM check */
int cpg_verifier_set_streams(void* verifier, int nstreams);
/* Cross-proof aggregation: `group` (a power of two, default 1 = off) consecutive proofs
 * are accepted by ONE MSM over their group*NV variable bases and ONE fixed-base MSM over their summed
 * CRS coefficients.  (Each proof's batching weights are already independent and secret-keyed, so the
 * sum of the proofs' relations is a random linear combination of all their checks.)  Groups that fail
 * are re-checked proof by proof: verdicts stay per proof and exact.
 * group = 0: adaptive - the library re-picks the group size after every batch from the observed rate
 * of failing proofs (cpg_verifier_group reads the current size).  group_window = 0 picks the window of
 * the aggregated MSM from its size.  cpg_verifier_rechecked = proofs of the last batch that took the
 * per-proof fallback. */
int cpg_verifier_set_group(void* verifier, int group, int group_window);
size_t cpg_verifier_rechecked(const void* verifier);
int cpg_verifier_group(const void* verifier);
/* Device-resident cache of decompressed tracker points, keyed by their 48-byte encodings (default off).  In Whisk the
 * pre-shuffle trackers of one shuffle proof are post-shuffle trackers of an earlier one (cp/whisk_interface.py:96-100
 * decodes all 4*ell of them on every call): a verifier that has met a tracker already knows its square root, and a
 * tracker that occurs several times in one batch is decompressed once.  Only the 4*ell tracker points of a proof go
 * through the table (M and the proof's own points are unique per proof).  2^log2_slots entries of 160 bytes (10..28;
 * 0 = off); open addressing, never evicts: the table starts over once it is half full.  Verdicts do not depend on it.
 * cache_stats: out3 = {lookups, lookups served from the table, slots claimed} since creation / cache_reset. */
int cpg_verifier_set_cache(void* verifier, int log2_slots);
int cpg_verifier_cache_reset(void* verifier);
int cpg_verifier_cache_stats(void* verifier, uint64_t* out3);
int cpg_verify_batch(void* verifier, const uint8_t* inputs, const uint8_t* proofs, size_t B, uint8_t* verdicts);
/* re-run the device side (decompress, D/A', MSM, test) of the last batch on its resident inputs */
int cpg_verify_replay_device(void* verifier, uint8_t* verdicts_or_null);

/* ---- batched shuffle-proof generation ----------------------------------------------------------
 * Replaces, for B proofs in lock-step, GenerateWhiskShuffleProof (cp/whisk_interface.py:111-140) ->
 * shuffle_permute_and_commit_input (cp/curdleproofs.py:301-321) + CurdleProofsProof.new (:50-160) and
 * every sub-proof's .new below it.  Transcript, Fr algebra and all group operations run on the GPU;
 * randomness is an input so that a caller drawing it in the reference's order gets the
 * reference's proof bytes.
 *   inputs    : [B][2*ell*48]  vec_R | vec_S (pre-shuffle tracker halves)
 *   perms     : [B][ell] u32   post[j] = k * pre[perm[j]]
 *   ks        : [B][32]        shuffle scalar k (canonical LE)
 *   rand      : [B][cpg_prover_rand_scalars][32]
 *               m_bl(4) a_bl(2) c_bl(4) ipa_r(n) ipa_z(n-2) r_t r_u r_a r_b r_k msm_r(n), canonical LE
 *   out_tu    : [B][2*ell*48]  vec_T | vec_U (post-shuffle tracker halves)
 *   out_proofs: [B][cpg_prover_proof_bytes]  M | proof  (WhiskShuffleProof.to_bytes, :57-61)
 *   status    : [B] 0 ok, 1 malformed input: a point encoding, k or a blinder >= r, or a perms row that is not a
 *               permutation of [0, ell).  Such a lane never reaches the device with its bad indices / scalars (it
 *               proves over a harmless substitute) and its outputs are undefined; the other lanes are unaffected.
 * Preconditions the caller keeps: every buffer holds B full rows of the sizes above. */
void* cpg_prover_create(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window);
/* BASELINE config 5's prove side: ONE proof at a time over ALL ranks of the communicator.  Every rank calls this and
 * every later cpg_prove_batch with identical arguments and gets identical outputs: the proof's leaves are split over
 * the ranks, each round's partial sums (<= 16 Jacobian points per proof) cross by ONE all-gather inside the library;
 * the transcript and the Fr vector kernels are replicated (cheap and deterministic). */
void* cpg_prover_create_sharded(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window);
int cpg_prover_free(void* prover);
size_t cpg_prover_proof_bytes(const void* prover);
size_t cpg_prover_rand_scalars(const void* prover);
int cpg_prover_set_window(void* prover, int var_window);
/* The post-shuffle trackers T_i, U_i each enter 8 of a proof's small MSMs: with window > 0 (default 6) the
 * prover builds, per batch, a table of the 2^(window-1) multiples of each of them and evaluates those MSMs
 * as table look-ups + Horner instead of the bucket method; 0 = bucket method for every variable-base MSM */
int cpg_prover_set_table_window(void* prover, int window);
/* sub-batches ("lanes", 1..8, default 2) whose rounds are issued alternately on separate streams: the
 * one-thread-per-proof transcript kernels of one lane run under the MSM kernels of the other */
int cpg_prover_set_lanes(void* prover, int nlanes, size_t min_proofs_per_lane /* 0 = 256: smaller batches are not split */);
/* where a proof's Fiat-Shamir transcript runs (its Fr vector work is always GPU kernels, one thread per element):
 * 0 = host threads (the reference's placement; right for a few large proofs: a CPU core runs the sequential Keccak
 * chain ~20x faster than one GPU thread), 1 = one GPU thread per proof (right for thousands of proofs: nothing
 * crosses PCIe between rounds), 2 (default) = by batch size */
int cpg_prover_set_transcript(void* prover, int mode);
int cpg_prove_replay_device(void* prover);   /* device side of the last batch again, inputs resident */
int cpg_prove_batch(void* prover, const uint8_t* inputs, const uint32_t* perms, const uint8_t* ks, const uint8_t* rand,
                    size_t B, uint8_t* out_tu, uint8_t* out_proofs, uint8_t* status);

/* ---- the prover's randomness, drawn as the reference draws it ----------------------
 * The reference takes the permutation (random.shuffle, cp/whisk_interface.py:114), k (:116) and every
 * blinder (random_scalar, cp/util.py:21-24) from Python's `random`.  This continues that very stream in C:
 * state625 = the 625 words of random.getstate()[1], advanced in place (random.setstate resumes from it).
 * For each of the B proofs: perm[ell], k, then n_rand = cpg_prover_rand_scalars blinders - exactly the
 * inputs of cpg_prove_batch.  Host-only; needs no device. */
int cpg_pyrandom_draw_shuffles(uint32_t* state625, size_t ell, size_t n_rand, size_t B,
                               uint32_t* perms, uint8_t* ks, uint8_t* rand);

/* ---- roofline support: saturating integer-pipe microbenchmark ---------------------------------
 * Runs `iters` dependent-chain steps of 32x32->64 multiply-accumulates on every SM and reports
 * the achieved rate: kind 0 = 32x32->64 MAC/s of data-dependent IMAD.WIDE.U32 chains (THE roofline
 * denominator: 32 lanes/clk/SM on B200), kind 1 = 32-bit IMAD/s (full-rate, context only),
 * kind 2 = Fq Montgomery products/s of this library's own field code. */
int cpg_bench_int_pipe(int kind, uint64_t iters, double* per_second, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* CPG_H */

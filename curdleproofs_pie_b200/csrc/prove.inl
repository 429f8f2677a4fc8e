// Batched shuffle-proof generation: B independent Whisk-size proofs in lock-step.
// Included at the end of cpg_api.cu after verify.inl (same translation unit).
//
// Replaces, for a whole batch at once, the reference's per-proof path
//   GenerateWhiskShuffleProof            curdleproofs/curdleproofs/whisk_interface.py:111-140
//   shuffle_permute_and_commit_input     curdleproofs/curdleproofs/curdleproofs.py:301-321
//   CurdleProofsProof.new                curdleproofs/curdleproofs/curdleproofs.py:50-160
//   SamePermutationProof.new / GrandProductProof.new / IPA.new / SameScalarProof.new / SameMSMProof.new
//                                        same_perm.py:27-72, grand_prod.py:29-119, ipa.py:27-48,75-153,
//                                        same_scalar.py:24-69, same_msm.py:50-144
// How.  Every group element the prover outputs is a linear combination of "leaves": the CRS points
// (vec_G | vec_H | H | G_t | G_u - evaluated through the fixed-base tables) and one of the per-proof
// vectors vec_R / vec_S / vec_T / vec_U (evaluated by the batched bucket method).  The reference's
// folded generator vectors G, G', T, U (ipa.py:145-146, same_msm.py:124-126) never materialise: a
// fold only rescales per-leaf weights, which is Fr work.  So a proof is 21 "rounds" of
//   step kernel (one proof per thread: absorb the previous round's points into the transcript, draw
//   challenges, update the Fr vectors, write the coefficient rows of this round's outputs)
//   -> fixed-base MSMs + batched Pippenger -> add -> compress
// and nothing crosses PCIe between the rounds.  Randomness is an INPUT (the caller draws it in the
// reference's order), so for a fixed Python `random` seed the proof bytes equal the
// reference's.  The prover's self-check asserts (grand_prod.py:103-105) are not replayed.
namespace {

// output ids (NOUT = 21 + 10 lg)
struct POut {
    uint32_t lg;
    uint32_t M = 0, A = 1, B = 2, C = 3, D = 4, Bc = 5, Bd = 6, ipa0 = 7;   // ipa0 + 4j + {L_C, L_D, R_C, R_D}
    uint32_t Rp, Sp, T1, T2, U1, U2, A1, A2, B1, B2, Ap, Ba, Bt, Bu, msm0, NOUT;  // msm0 + 6j + {L_A, L_T, L_U, R_A, R_T, R_U}
    explicit POut(uint32_t lg_) : lg(lg_) {
        Rp = ipa0 + 4 * lg; Sp = Rp + 1; T1 = Rp + 2; T2 = Rp + 3; U1 = Rp + 4; U2 = Rp + 5; A1 = Rp + 6; A2 = Rp + 7; B1 = Rp + 8; B2 = Rp + 9;
        Ap = Rp + 10; Ba = Ap + 1; Bt = Ap + 2; Bu = Ap + 3; msm0 = Ap + 4; NOUT = msm0 + 6 * lg;
    }
};

constexpr uint32_t P_MAX_OUT = 10;     // outputs per round (the SameScalar round has 10)
constexpr uint32_t P_MAX_VAR = 6;      // of which with a variable-base part

struct PShape { uint32_t ell, n, lg, NF, NR, rounds; uint32_t CH, nch; };   // CH * nch >= n: chunks of the two-pass reductions

// A round of a proof is split in two:
//   * the TRANSCRIPT step (prove_transcript): strictly sequential Keccak work plus O(1) scalars - absorb the previous
//     round's outputs, squeeze this round's challenges.  One proof per GPU thread for big batches, or on host threads
//     for a few (large) proofs, where a CPU core runs the Keccak chain ~20x faster than a lone GPU thread (the north
//     star's placement: "the sequential Merlin/STROBE transcript stays on the host").  Same CPG_HD source either way.
//   * the VECTOR kernels (ProveVec, ProveReduce*, ProveScan*): all Fr vector work - the grand product and its prefix
//     products, the blinding polynomials, the IPA / SameMSM folds and the coefficient rows of the round's MSMs - one
//     thread per (proof, element), sums and products in two short passes (cp/grand_prod.py:49-51, cp/ipa.py:33-46,
//     142-146, cp/same_msm.py:122-126).  A lone n = 16384 proof therefore runs 16 384 threads wide instead of one.
// The three small structs below are all that crosses between the two halves (and, with the transcript on the host, PCIe).
struct PChal {                    // transcript -> vector kernels: the challenges in force
    HFr alpha_sp, beta_sp, alpha_gp, beta_gp, beta_gp_inv, alpha_ipa, beta_ipa, gam, gam_inv, alpha_msm;
};
struct PRes {                     // vector kernels -> transcript: scalars the transcript absorbs or the wire proof carries
    HFr gprod, r_p, z, c0, c1, d0, d1, x0, x1;
};
struct PTr {                      // transcript-side state of one proof
    cpgh::Transcript tr;
    HFr k, c_final, d_final, z_k, z_t, z_u;
};

// vectors per proof, each n entries: see PV_* below
enum { PV_APERM = 0, PV_FACT, PV_C, PV_D, PV_X, PV_WG, PV_WGP, PV_W2, PV_ACOEF, PV_MCOEF, PV_BCOEF, PV_U, PV_COUNT };

struct PBuffers {                 // device side (the vector kernels, and the transcript when it runs on the device)
    const uint8_t* in48;          // [B][2 ell][48]  vec_R | vec_S wire bytes
    const uint8_t* tu48;          // [B][2 ell][48]  vec_T | vec_U wire bytes (round 0 output)
    const uint32_t* perm;         // [B][ell]
    const uint8_t* kbytes;        // [B][32]
    const uint8_t* rand;          // [B][NR][32]     m_bl(4) a_bl(2) c_bl(4) ipa_r(n) ipa_z(n-2) r_t r_u r_a r_b r_k msm_r(n)
    const uint8_t* crs48;         // CRS wire bytes
    PTr* trs;                     // [B]
    PChal* chal;                  // [B]
    PRes* res;                    // [B]
    HFr* achal;                   // [B][ell]        the vec_a challenges, proof-major (written by the transcript)
    HFr* vec;                     // [PV_COUNT][n][B]  (batch-innermost, see FrVec)
    HFr* part;                    // [2][nch][B]     partial sums / products of the two-pass reductions
    HFr* red;                     // [2][B]          their totals
    uint8_t* outs48;              // [B][NOUT][48]   compressed outputs, by output id
    uint8_t* fs;                  // [P_MAX_OUT][B][NF][32]  fixed-base coefficient rows of the round being prepared (output-major)
    uint8_t* vs;                  // [P_MAX_VAR][B][ell][32] variable-base coefficient rows
    uint8_t* proof;               // [B][proof_len]  wire proof (without M), filled at the end
    uint64_t B;
};

// A per-proof Fr vector.  Storage is batch-innermost, vec[which][i][b]: consecutive threads (consecutive proofs, same
// element) touch consecutive words.
struct FrVec {
    HFr* p; uint64_t stride;
    CPG_HD HFr& operator[](uint32_t i) const { return p[(uint64_t)i * stride]; }
    CPG_HD FrVec operator+(uint32_t k) const { return FrVec{p + (uint64_t)k * stride, stride}; }
};
CPG_HD FrVec pvec(const PShape& sh, const PBuffers& pb, size_t b, int which) { return FrVec{pb.vec + (size_t)which * sh.n * pb.B + b, pb.B}; }
CPG_HD uint8_t* frow(const PShape& sh, const PBuffers& pb, size_t b, uint32_t o) { return pb.fs + ((size_t)o * pb.B + b) * (size_t)sh.NF * 32; }
CPG_HD uint8_t* vrow(const PShape& sh, const PBuffers& pb, size_t b, uint32_t o) { return pb.vs + ((size_t)o * pb.B + b) * (size_t)sh.ell * 32; }
CPG_HD HFr prand_at(const uint8_t* rand, const PShape& sh, size_t b, uint32_t i) {
    HFr v; cpgh::fr_from_bytes(&v, rand + (b * sh.NR + i) * 32); return v;   // canonical: cpg_prove_batch checked every blinder
}
CPG_HD HFr prand(const PBuffers& pb, const PShape& sh, size_t b, uint32_t i) { return prand_at(pb.rand, sh, b, i); }
CPG_HD void put(uint8_t* row, uint32_t i, const HFr& v) { cpgh::fr_to_bytes(row + 32 * (size_t)i, v); }
CPG_HD void put0(uint8_t* row, uint32_t i) { uint64_t* q = (uint64_t*)(row + 32 * (size_t)i); q[0] = q[1] = q[2] = q[3] = 0; }
// fixed-table index of leaf L of G_wb = vec_G | vec_H[:2] | G_t | G_u, and back (-1: the table entry is not a leaf of G_wb)
CPG_HD uint32_t gwb_index(const PShape& sh, uint32_t L) { return L < sh.ell + 2 ? L : L + 3; }   // ell+2 -> n+1 (G_t), ell+3 -> n+2 (G_u)
CPG_HD int gwb_leaf(const PShape& sh, uint32_t i) { return i < sh.ell + 2 ? (int)i : (i == sh.n + 1 ? (int)sh.ell + 2 : (i == sh.n + 2 ? (int)sh.ell + 3 : -1)); }

// offsets into the rand array
struct PRand { uint32_t m_bl = 0, a_bl = 4, c_bl = 6, ipa_r = 10, ipa_z, r_t, r_u, r_a, r_b, r_k, msm_r, NR;
    CPG_HD explicit PRand(uint32_t n) { ipa_z = ipa_r + n; r_t = ipa_z + (n - 2); r_u = r_t + 1; r_a = r_u + 1; r_b = r_a + 1; r_k = r_b + 1; msm_r = r_k + 1; NR = msm_r + n; } };

// Rounds: 0 M | 1 A | 2 B | 3 C | 4 D,B_c,B_d | 5..4+lg IPA | 5+lg SameScalar | 6+lg A',B_a,B_t,B_u |
// 7+lg..6+2lg SameMSM | 7+2lg finish (assemble the wire proof).
struct PRounds { uint32_t IPA0, SS, MSM_INIT, MSM0, FIN;
    CPG_HD explicit PRounds(uint32_t lg) { IPA0 = 5; SS = 5 + lg; MSM_INIT = 6 + lg; MSM0 = 7 + lg; FIN = 7 + 2 * lg; } };

// ---- the transcript step of one round of one proof -------------------------------------------------------------------
// Every pointer lives in the memory space of whoever runs it (device buffers, or the host mirrors of ProverLane).
struct PTBuf {
    const uint8_t* in48; const uint8_t* tu48; const uint8_t* kbytes; const uint8_t* rand; const uint8_t* crs48;
    const uint8_t* outs48;        // [B][NOUT][48]
    PTr* trs; PChal* chal; const PRes* res; HFr* achal;
    uint8_t* proof;               // [B][proof_len]
};
CPG_HD void prove_transcript(const PShape& sh, const POut& O, const PTBuf& pt, uint32_t round, size_t b) {
    using namespace cpgh;
    const uint32_t ell = sh.ell, n = sh.n, lg = sh.lg;
    const PRand RO(n);
    const PRounds R(lg);
    if (round == 0) return;                                                // M needs no challenge
    PTr& s = pt.trs[b];
    PChal ch = pt.chal[b];
    const PRes& rs = pt.res[b];
    const uint8_t* outs = pt.outs48 + b * (size_t)O.NOUT * 48;
    HFr* a = pt.achal + b * (size_t)ell;
    // the STROBE state is worked on in thread-local storage and written back once (per-thread structs in global memory
    // make every byte XOR an uncoalesced read-modify-write)
    Transcript tr;
    if (round > 1) tr = s.tr;
    if (round == 1) {                                   // step1 -> vec_a                              (curdleproofs.py:65-71)
        fr_from_bytes(&s.k, pt.kbytes + b * 32);
        tr.init("curdleproofs");
        const uint8_t* in = pt.in48 + b * (size_t)(2 * ell) * 48;
        const uint8_t* tu = pt.tu48 + b * (size_t)(2 * ell) * 48;
        for (uint32_t i = 0; i < 2 * ell; i++) tr.append_point("curdleproofs_step1", in + 48 * (size_t)i);
        for (uint32_t i = 0; i < 2 * ell; i++) tr.append_point("curdleproofs_step1", tu + 48 * (size_t)i);
        tr.append_point("curdleproofs_step1", outs + 48 * O.M);
        for (uint32_t i = 0; i < ell; i++) a[i] = tr.challenge("curdleproofs_vec_a");
    } else if (round == 2) {                            // same_perm: alpha, beta                     (same_perm.py:43-46)
        tr.append_point("same_perm_step1", outs + 48 * O.A);
        tr.append_point("same_perm_step1", outs + 48 * O.M);
        for (uint32_t i = 0; i < ell; i++) tr.append_fr("same_perm_step1", a[i]);
        ch.alpha_sp = tr.challenge("same_perm_alpha");
        ch.beta_sp = tr.challenge("same_perm_beta");
    } else if (round == 3) {                            // gprod step 1: alpha                        (grand_prod.py:44-46)
        tr.append_point("gprod_step1", outs + 48 * O.B);
        tr.append_fr("gprod_step1", rs.gprod);
        ch.alpha_gp = tr.challenge("gprod_alpha");
    } else if (round == 4) {                            // gprod step 2: beta                         (grand_prod.py:59-61)
        tr.append_point("gprod_step2", outs + 48 * O.C);
        tr.append_fr("gprod_step2", rs.r_p);
        ch.beta_gp = tr.challenge("gprod_beta");
        ch.beta_gp_inv = fr_inv(ch.beta_gp);
    } else if (round == R.IPA0) {                       // IPA: alpha, beta                           (ipa.py:100-105)
        tr.append_point("ipa_step1", outs + 48 * O.C);
        tr.append_point("ipa_step1", outs + 48 * O.D);
        tr.append_fr("ipa_step1", rs.z);
        tr.append_point("ipa_step1", outs + 48 * O.Bc);
        tr.append_point("ipa_step1", outs + 48 * O.Bd);
        ch.alpha_ipa = tr.challenge("ipa_alpha");
        ch.beta_ipa = tr.challenge("ipa_beta");
    } else if (round > R.IPA0 && round <= R.SS) {       // IPA rounds: gamma                          (ipa.py:136-139)
        const uint8_t* pr = outs + 48 * (size_t)(O.ipa0 + 4 * (round - R.IPA0 - 1));
        for (uint32_t k = 0; k < 4; k++) tr.append_point("ipa_loop", pr + 48 * k);
        ch.gam = tr.challenge("ipa_gamma");
        ch.gam_inv = fr_inv(ch.gam);
        if (round == R.SS) {                            // last fold: c_final, d_final
            s.c_final = fr_add(rs.c0, fr_mul(ch.gam_inv, rs.c1));
            s.d_final = fr_add(rs.d0, fr_mul(ch.gam, rs.d1));
        }
    } else if (round == R.MSM_INIT) {                   // SameScalar: alpha, responses               (same_scalar.py:46-63)
        const uint32_t ss[10] = {O.Rp, O.Sp, O.T1, O.T2, O.U1, O.U2, O.A1, O.A2, O.B1, O.B2};
        for (uint32_t k = 0; k < 10; k++) tr.append_point("sameexp_points", outs + 48 * (size_t)ss[k]);
        HFr alpha = tr.challenge("same_scalar_alpha");
        const HFr r_t = prand_at(pt.rand, sh, b, RO.r_t), r_u = prand_at(pt.rand, sh, b, RO.r_u);
        const HFr r_a = prand_at(pt.rand, sh, b, RO.r_a), r_b = prand_at(pt.rand, sh, b, RO.r_b), r_k = prand_at(pt.rand, sh, b, RO.r_k);
        s.z_k = fr_add(r_k, fr_mul(s.k, alpha));
        s.z_t = fr_add(r_a, fr_mul(r_t, alpha));
        s.z_u = fr_add(r_b, fr_mul(r_u, alpha));
    } else if (round == R.MSM0) {                       // SameMSM: alpha                             (same_msm.py:79-88)
        tr.append_point("same_msm_step1", outs + 48 * O.Ap);
        tr.append_point("same_msm_step1", outs + 48 * O.T2);
        tr.append_point("same_msm_step1", outs + 48 * O.U2);
        const uint8_t* INF = CPGH_SEL(INF48);
        const uint8_t* Hb = pt.crs48 + 48 * (size_t)n;
        const uint8_t* tu = pt.tu48 + b * (size_t)(2 * ell) * 48;
        for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", tu + 48 * (size_t)i);
        tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
        tr.append_point("same_msm_step1", Hb); tr.append_point("same_msm_step1", INF);
        for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", tu + 48 * (size_t)(ell + i));
        tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
        tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", Hb);
        tr.append_point("same_msm_step1", outs + 48 * O.Ba);
        tr.append_point("same_msm_step1", outs + 48 * O.Bt);
        tr.append_point("same_msm_step1", outs + 48 * O.Bu);
        ch.alpha_msm = tr.challenge("same_msm_alpha");
    } else if (round > R.MSM0 && round <= R.FIN) {      // SameMSM rounds: gamma                      (same_msm.py:115-119)
        const uint8_t* pr = outs + 48 * (size_t)(O.msm0 + 6 * (round - R.MSM0 - 1));
        for (uint32_t k = 0; k < 6; k++) tr.append_point("same_msm_loop", pr + 48 * k);
        ch.gam = tr.challenge("same_msm_gamma");
        ch.gam_inv = fr_inv(ch.gam);
        if (round == R.FIN) {                           // last fold + wire assembly                  (curdleproofs.py:275-285)
            HFr x_final = fr_add(rs.x0, fr_mul(ch.gam_inv, rs.x1));
            uint8_t* w = pt.proof + b * (size_t)(1088 + 480 * (size_t)lg);
            auto ptc = [&](uint32_t id) { memcpy(w, outs + 48 * (size_t)id, 48); w += 48; };
            auto sc = [&](const HFr& v) { fr_to_bytes(w, v); w += 32; };
            ptc(O.A); ptc(O.T1); ptc(O.T2); ptc(O.U1); ptc(O.U2); ptc(O.Rp); ptc(O.Sp);
            ptc(O.B); ptc(O.C); sc(rs.r_p);
            ptc(O.Bc); ptc(O.Bd);
            for (uint32_t k = 0; k < 4; k++) {           // L_C[], R_C[], L_D[], R_D[]  (ipa.py:260-270)
                const uint32_t sel[4] = {0, 2, 1, 3};    // stored per round as L_C, L_D, R_C, R_D
                for (uint32_t j = 0; j < lg; j++) ptc(O.ipa0 + 4 * j + sel[k]);
            }
            sc(s.c_final); sc(s.d_final);
            ptc(O.A1); ptc(O.A2); ptc(O.B1); ptc(O.B2); sc(s.z_k); sc(s.z_t); sc(s.z_u);
            ptc(O.Ba); ptc(O.Bt); ptc(O.Bu);
            for (uint32_t k = 0; k < 6; k++) for (uint32_t j = 0; j < lg; j++) ptc(O.msm0 + 6 * j + k);   // L_A L_T L_U R_A R_T R_U
            sc(x_final);
        }
    }
    s.tr = tr;
    pt.chal[b] = ch;
}
struct ProveTranscript {          // thread = proof (device placement)
    static constexpr const char* kName = "ProveTranscript";
    PShape sh; POut O; PTBuf pt; uint32_t round;
    CPG_HD void operator()(uint64_t b) const { prove_transcript(sh, O, pt, round, (size_t)b); }
};

// ---- the vector kernels -------------------------------------------------------------------------------------------
// Element-wise stage `phase` of round `round` for element i of proof b; i runs over the NF = n + 3 entries of a
// fixed-base coefficient row (vec_G | vec_H | H | G_t | G_u), of which the first n are also the indices of the Fr vectors.
// Every entry of every coefficient row of the round is written (zeros included), so rows need no separate clearing.
CPG_HD void prove_vec(const PShape& sh, const PBuffers& pb, uint32_t round, uint32_t phase, size_t b, uint32_t i) {
    using namespace cpgh;
    const uint32_t ell = sh.ell, n = sh.n, lg = sh.lg;
    const PRand RO(n);
    const PRounds R(lg);
    const PChal& ch = pb.chal[b];
    PRes& rs = pb.res[b];
    const uint32_t* perm = pb.perm + b * (size_t)ell;
    FrVec aperm = pvec(sh, pb, b, PV_APERM), fact = pvec(sh, pb, b, PV_FACT), c = pvec(sh, pb, b, PV_C), d = pvec(sh, pb, b, PV_D);
    FrVec x = pvec(sh, pb, b, PV_X), wG = pvec(sh, pb, b, PV_WG), wGp = pvec(sh, pb, b, PV_WGP), w2 = pvec(sh, pb, b, PV_W2);
    FrVec Acoef = pvec(sh, pb, b, PV_ACOEF), Mcoef = pvec(sh, pb, b, PV_MCOEF), Bcoef = pvec(sh, pb, b, PV_BCOEF), uvec = pvec(sh, pb, b, PV_U);
    uint8_t* f0 = frow(sh, pb, b, 0);

    if (round == 0) {                                   // M = MSM(vec_G, sigma) + MSM(vec_H, m_bl)   (curdleproofs.py:310-319)
        if (i < n) { HFr v = i < ell ? fr_from_u64(perm[i]) : prand(pb, sh, b, RO.m_bl + (i - ell)); Mcoef[i] = v; put(f0, i, v); }
        else put0(f0, i);
        return;
    }
    if (round == 1) {                                   // A = MSM(vec_G, sigma(a)) + MSM(vec_H, a_bl | 0 0)  (curdleproofs.py:72-77)
        if (i < n) {
            HFr v = i < ell ? pb.achal[b * (size_t)ell + perm[i]] : (i < ell + 2 ? prand(pb, sh, b, RO.a_bl + (i - ell)) : fr_zero());
            aperm[i] = v; Acoef[i] = v; put(f0, i, v);
        } else put0(f0, i);
        return;
    }
    if (round == 2) {                                   // B; the grand product's factors              (same_perm.py:49-55)
        if (i < n) {
            if (i < ell) fact[i] = fr_add(fr_add(aperm[i], fr_mul(fr_from_u64(perm[i]), ch.alpha_sp)), ch.beta_sp);
            HFr t = fr_add(Acoef[i], fr_mul(ch.alpha_sp, Mcoef[i]));
            if (i < ell) t = fr_add(t, ch.beta_sp);
            Bcoef[i] = t; put(f0, i, t);
        } else put0(f0, i);
        return;
    }
    if (round == 3) {                                   // C (prefix products c[0..ell) come from ProveScan)   (grand_prod.py:49-57)
        if (i < ell) put(f0, i, c[i]);
        else if (i < n) {
            HFr cb = prand(pb, sh, b, RO.c_bl + (i - ell));
            c[i] = cb;
            d[i] = fr_add(fr_add(aperm[i], fr_mul(ch.alpha_sp, Mcoef[i])), ch.alpha_gp);   // rb_alpha, parked in d's blinder slots
            put(f0, i, cb);
        } else put0(f0, i);
        if (i == 0) {                                   // r_p = <rb_alpha, c_bl>
            HFr rp = fr_zero();
            for (uint32_t q = 0; q < 4; q++) {
                HFr rb = fr_add(fr_add(aperm[ell + q], fr_mul(ch.alpha_sp, Mcoef[ell + q])), ch.alpha_gp);
                rp = fr_add(rp, fr_mul(rb, prand(pb, sh, b, RO.c_bl + q)));
            }
            rs.r_p = rp;
        }
        return;
    }
    if (round == 4 && phase == 0) {                     // d, u = beta^-(i+1), IPA blinders            (grand_prod.py:66-89, ipa.py:27-31)
        if (i < n) {
            if (i < ell) {
                HFr pw = fr_pow_u64(ch.beta_gp, i), pw1 = fr_mul(pw, ch.beta_gp);
                d[i] = fr_sub(fr_mul(fact[i], pw1), pw);                 // b_i beta^(i+1) - beta^i
                uvec[i] = fr_pow_u64(ch.beta_gp_inv, i + 1);
            } else {
                d[i] = fr_mul(fr_pow_u64(ch.beta_gp, ell + 1), d[i]);   // d_bl = beta^(ell+1) rb_alpha
                uvec[i] = fr_pow_u64(ch.beta_gp_inv, ell + 1);
            }
            x[i] = prand(pb, sh, b, RO.ipa_r + i);                      // r (scratch in x)
            if (i + 2 < n) w2[i] = prand(pb, sh, b, RO.ipa_z + i);       // z (scratch in w2); its last two entries follow the sums
        }
        if (i == 0) {
            HFr beta_l = fr_pow_u64(ch.beta_gp, ell), beta_l1 = fr_mul(beta_l, ch.beta_gp);
            rs.z = fr_sub(fr_add(fr_mul(rs.r_p, beta_l1), fr_mul(rs.gprod, beta_l)), fr_one());
        }
        return;
    }
    if (round == 4) {                                   // rows of D, B_c, B_d                          (grand_prod.py:90, ipa.py:97-98)
        uint8_t *fD = f0, *fBc = frow(sh, pb, b, 1), *fBd = frow(sh, pb, b, 2);
        if (i < n) {
            HFr t = i < ell ? fr_sub(Bcoef[i], ch.beta_gp_inv) : fr_add(Bcoef[i], ch.alpha_gp);
            HFr r = x[i], zz = w2[i];
            put(fD, i, t); put(fBc, i, r); put(fBd, i, fr_mul(zz, uvec[i]));
            wG[i] = r; wGp[i] = zz;                                       // r_c / r_d wait here until alpha is known
        } else { put0(fD, i); put0(fBc, i); put0(fBd, i); }
        return;
    }
    if (round >= R.IPA0 && round < R.SS) {              // IPA rounds                                   (ipa.py:107-151)
        const uint32_t j = round - R.IPA0, len = n >> j, m = len / 2;
        if (phase == 0) {
            if (j == 0) {
                if (i < n) {                            // c = r_c + alpha c ; d = r_d + alpha d ; leaf weights reset
                    c[i] = fr_add(wG[i], fr_mul(ch.alpha_ipa, c[i]));
                    d[i] = fr_add(wGp[i], fr_mul(ch.alpha_ipa, d[i]));
                    wG[i] = fr_one(); wGp[i] = uvec[i];
                }
            } else {                                    // fold the halves of the previous length 2 len
                if (i < len) { c[i] = fr_add(c[i], fr_mul(ch.gam_inv, c[len + i])); d[i] = fr_add(d[i], fr_mul(ch.gam, d[len + i])); }
                if (i < n && ((i / len) & 1)) { wG[i] = fr_mul(wG[i], ch.gam); wGp[i] = fr_mul(wGp[i], ch.gam_inv); }
            }
            return;
        }
        uint8_t *fLC = f0, *fLD = frow(sh, pb, b, 1), *fRC = frow(sh, pb, b, 2), *fRD = frow(sh, pb, b, 3);
        if (i < n) {
            const uint32_t q = i % m;
            if ((i / m) & 1) {                          // the leaf folds into the right half
                put(fLC, i, fr_mul(c[q], wG[i])); put(fRD, i, fr_mul(d[q], wGp[i]));           // MSM(G_R, c_L), MSM(G'_R, d_L)
                put0(fRC, i); put0(fLD, i);
            } else {
                put(fRC, i, fr_mul(c[m + q], wG[i])); put(fLD, i, fr_mul(d[m + q], wGp[i]));   // MSM(G_L, c_R), MSM(G'_L, d_R)
                put0(fLC, i); put0(fRD, i);
            }
        } else if (i == n) {                            // + H beta <c_L, d_R> and + H beta <c_R, d_L>
            put(fLC, i, fr_mul(ch.beta_ipa, pb.red[b])); put(fRC, i, fr_mul(ch.beta_ipa, pb.red[pb.B + b]));
            put0(fLD, i); put0(fRD, i);
        } else { put0(fLC, i); put0(fLD, i); put0(fRC, i); put0(fRD, i); }
        if (i == 0 && len == 2) { rs.c0 = c[0]; rs.c1 = c[1]; rs.d0 = d[0]; rs.d1 = d[1]; }
        return;
    }
    if (round == R.SS) {                                // R', S', cm_T, cm_U, cm_A, cm_B              (curdleproofs.py:92-102, same_scalar.py:39-44)
        // outputs: Rp Sp T1 T2 U1 U2 A1 A2 B1 B2 ; var rows: 0 R' = MSM(vec_R, a), 1 S' = MSM(vec_S, a); the commitments'
        // k R', r_k R', k S', r_k S' are one scalar-mul each of those results (ProveCombine)
        for (uint32_t o = 0; o < 10; o++) put0(frow(sh, pb, b, o), i);
        if (i == n) {                                   // H
            put(frow(sh, pb, b, 3), i, prand(pb, sh, b, RO.r_t)); put(frow(sh, pb, b, 5), i, prand(pb, sh, b, RO.r_u));
            put(frow(sh, pb, b, 7), i, prand(pb, sh, b, RO.r_a)); put(frow(sh, pb, b, 9), i, prand(pb, sh, b, RO.r_b));
        } else if (i == n + 1) {                        // G_t
            put(frow(sh, pb, b, 2), i, prand(pb, sh, b, RO.r_t)); put(frow(sh, pb, b, 6), i, prand(pb, sh, b, RO.r_a));
        } else if (i == n + 2) {                        // G_u
            put(frow(sh, pb, b, 4), i, prand(pb, sh, b, RO.r_u)); put(frow(sh, pb, b, 8), i, prand(pb, sh, b, RO.r_b));
        }
        if (i < ell) { uint8_t* vR = vrow(sh, pb, b, 0); put(vR, i, pb.achal[b * (size_t)ell + i]); cpgh::copy32(vrow(sh, pb, b, 1) + 32 * (size_t)i, vR + 32 * (size_t)i); }
        return;
    }
    if (round == R.MSM_INIT) {                          // A', B_a, B_t, B_u                             (same_msm.py:73-77)
        // x_wb = a_perm | a_bl | r_t r_u ; the blinders r wait in w2 until alpha_msm is known
        if (i < n) {
            x[i] = i < ell + 2 ? aperm[i] : prand(pb, sh, b, i == ell + 2 ? RO.r_t : RO.r_u);
            w2[i] = prand(pb, sh, b, RO.msm_r + i);
        }
        uint8_t *fAp = f0, *fBa = frow(sh, pb, b, 1), *fBt = frow(sh, pb, b, 2), *fBu = frow(sh, pb, b, 3);
        if (i < n) put(fAp, i, Acoef[i]);
        else if (i == n) put0(fAp, i);
        else put(fAp, i, prand(pb, sh, b, i == n + 1 ? RO.r_t : RO.r_u));      // + cm_T.T_1 = G_t r_t, + cm_U.T_1 = G_u r_u
        const int L = gwb_leaf(sh, i);
        if (L >= 0) put(fBa, i, prand(pb, sh, b, RO.msm_r + (uint32_t)L)); else put0(fBa, i);
        if (i == n) { put(fBt, i, prand(pb, sh, b, RO.msm_r + ell + 2)); put(fBu, i, prand(pb, sh, b, RO.msm_r + ell + 3)); }   // T_wb = vec_T | 0 0 H 0, U_wb = vec_U | 0 0 0 H
        else { put0(fBt, i); put0(fBu, i); }
        if (i < ell) { uint8_t* vT = vrow(sh, pb, b, 0); put(vT, i, prand(pb, sh, b, RO.msm_r + i)); cpgh::copy32(vrow(sh, pb, b, 1) + 32 * (size_t)i, vT + 32 * (size_t)i); }
        return;
    }
    if (round >= R.MSM0 && round < R.FIN) {             // SameMSM rounds                                 (same_msm.py:90-131)
        const uint32_t j = round - R.MSM0, len = n >> j, m = len / 2;
        if (phase == 0) {
            if (j == 0) { if (i < n) { x[i] = fr_add(w2[i], fr_mul(ch.alpha_msm, x[i])); w2[i] = fr_one(); } }
            else {
                if (i < len) x[i] = fr_add(x[i], fr_mul(ch.gam_inv, x[len + i]));
                if (i < n && ((i / len) & 1)) w2[i] = fr_mul(w2[i], ch.gam);
            }
            return;
        }
        // outputs: 0 L_A 1 L_T 2 L_U 3 R_A 4 R_T 5 R_U ; var rows: 0 L_T (T) 1 L_U (U) 2 R_T (T) 3 R_U (U)
        // L_* = MSM(v[m:], x_L), R_* = MSM(v[:m], x_R): a leaf in a right half carries x[q] into L_*, else x[m + q] into R_*
        uint8_t *fLA = f0, *fLT = frow(sh, pb, b, 1), *fLU = frow(sh, pb, b, 2), *fRA = frow(sh, pb, b, 3), *fRT = frow(sh, pb, b, 4), *fRU = frow(sh, pb, b, 5);
        auto coef_of = [&](uint32_t L, bool* right) { const uint32_t q = L % m; *right = ((L / m) & 1) != 0; return fr_mul(*right ? x[q] : x[m + q], w2[L]); };
        const int L = gwb_leaf(sh, i);
        if (L >= 0) { bool right; HFr cf = coef_of((uint32_t)L, &right); if (right) { put(fLA, i, cf); put0(fRA, i); } else { put(fRA, i, cf); put0(fLA, i); } }
        else { put0(fLA, i); put0(fRA, i); }
        put0(fLT, i); put0(fLU, i); put0(fRT, i); put0(fRU, i);
        if (i == n) {                                   // H inside T_wb (leaf ell + 2) and inside U_wb (leaf ell + 3)
            bool right; HFr cf = coef_of(ell + 2, &right); put(right ? fLT : fRT, i, cf);
            cf = coef_of(ell + 3, &right); put(right ? fLU : fRU, i, cf);
        }
        if (i < ell) {
            bool right; HFr cf = coef_of(i, &right);
            uint8_t *vLT = vrow(sh, pb, b, 0), *vLU = vrow(sh, pb, b, 1), *vRT = vrow(sh, pb, b, 2), *vRU = vrow(sh, pb, b, 3);
            if (right) { put(vLT, i, cf); cpgh::copy32(vLU + 32 * (size_t)i, vLT + 32 * (size_t)i); put0(vRT, i); put0(vRU, i); }
            else { put(vRT, i, cf); cpgh::copy32(vRU + 32 * (size_t)i, vRT + 32 * (size_t)i); put0(vLT, i); put0(vLU, i); }
        }
        if (i == 0 && len == 2) { rs.x0 = x[0]; rs.x1 = x[1]; }
        return;
    }
}
struct ProveVec {                 // thread = (element i, proof b), proofs innermost
    static constexpr const char* kName = "ProveVec";
    PShape sh; PBuffers pb; uint32_t round, phase;
    CPG_HD void operator()(uint64_t t) const { prove_vec(sh, pb, round, phase, (size_t)(t % pb.B), (uint32_t)(t / pb.B)); }
};

// Sums and products over a proof's vectors in two short passes: thread (chunk, b) reduces CH consecutive elements,
// thread b then combines the nch partials (serial depth CH + nch ~ 2 sqrt(n) instead of n).
enum { PR_PROD_FACT = 0, PR_IPA_BLIND = 1, PR_IPA_LR = 2 };
struct ProveReduce1 {
    static constexpr const char* kName = "ProveReduce1";
    PShape sh; PBuffers pb; uint32_t kind, m;             // m: half length of the IPA round (PR_IPA_LR)
    CPG_HD void operator()(uint64_t t) const {
        using namespace cpgh;
        const size_t b = (size_t)(t % pb.B); const uint32_t k = (uint32_t)(t / pb.B);
        const uint32_t lo = k * sh.CH, hi0 = lo + sh.CH;
        HFr* p0 = pb.part + (size_t)k * pb.B + b; HFr* p1 = pb.part + ((size_t)sh.nch + k) * pb.B + b;
        if (kind == PR_PROD_FACT) {                        // prod_{i < ell} fact[i]
            FrVec fact = pvec(sh, pb, b, PV_FACT);
            HFr acc = fr_one();
            for (uint32_t i = lo; i < hi0 && i < sh.ell; i++) acc = fr_mul(acc, fact[i]);
            *p0 = acc;
        } else if (kind == PR_IPA_BLIND) {                 // omega = <r, d> + <z, c>_{n-2}, delta = <r, z>_{n-2}   (ipa.py:33-35)
            FrVec r = pvec(sh, pb, b, PV_X), zz = pvec(sh, pb, b, PV_W2), c = pvec(sh, pb, b, PV_C), d = pvec(sh, pb, b, PV_D);
            HFr om = fr_zero(), de = fr_zero();
            for (uint32_t i = lo; i < hi0 && i < sh.n; i++) {
                om = fr_add(om, fr_mul(r[i], d[i]));
                if (i + 2 < sh.n) { om = fr_add(om, fr_mul(zz[i], c[i])); de = fr_add(de, fr_mul(r[i], zz[i])); }
            }
            *p0 = om; *p1 = de;
        } else {                                           // <c_L, d_R>, <c_R, d_L> over the current halves       (ipa.py:126-129)
            FrVec c = pvec(sh, pb, b, PV_C), d = pvec(sh, pb, b, PV_D);
            HFr l = fr_zero(), r = fr_zero();
            for (uint32_t i = lo; i < hi0 && i < m; i++) { l = fr_add(l, fr_mul(c[i], d[m + i])); r = fr_add(r, fr_mul(c[m + i], d[i])); }
            *p0 = l; *p1 = r;
        }
    }
};
struct ProveReduce2 {             // thread = proof
    static constexpr const char* kName = "ProveReduce2";
    PShape sh; PBuffers pb; uint32_t kind;
    CPG_HD void operator()(uint64_t b) const {
        using namespace cpgh;
        const HFr* p0 = pb.part + b; const HFr* p1 = pb.part + (size_t)sh.nch * pb.B + b;
        if (kind == PR_PROD_FACT) {
            HFr acc = fr_one();
            for (uint32_t k = 0; k < sh.nch; k++) acc = fr_mul(acc, p0[(size_t)k * pb.B]);
            pb.res[b].gprod = acc;
            return;
        }
        HFr s0 = fr_zero(), s1 = fr_zero();
        for (uint32_t k = 0; k < sh.nch; k++) { s0 = fr_add(s0, p0[(size_t)k * pb.B]); s1 = fr_add(s1, p1[(size_t)k * pb.B]); }
        if (kind == PR_IPA_LR) { pb.red[b] = s0; pb.red[pb.B + b] = s1; return; }
        // the last two entries of the blinder z make <r, d> + <z, c> = 0 and <r, z> = 0 (ipa.py:36-46)
        const uint32_t n = sh.n;
        FrVec r = pvec(sh, pb, (size_t)b, PV_X), zz = pvec(sh, pb, (size_t)b, PV_W2), c = pvec(sh, pb, (size_t)b, PV_C);
        const HFr omega = s0, delta = s1;
        HFr inv_c = fr_inv(c[n - 2]);
        HFr t1 = fr_mul(r[n - 2], inv_c);
        HFr last_z = fr_mul(fr_sub(fr_mul(t1, omega), delta), fr_inv(fr_add(fr_neg(fr_mul(t1, c[n - 1])), r[n - 1])));
        HFr pen_z = fr_neg(fr_mul(inv_c, fr_add(fr_mul(last_z, c[n - 1]), omega)));
        zz[n - 2] = pen_z; zz[n - 1] = last_z;
    }
};
// c[i] = prod_{q < i} fact[q], i < ell (cp/grand_prod.py:49-51, a sequential loop there): chunk products (ProveReduce1),
// an exclusive scan of the nch chunk products by one thread per proof, then every chunk expands its own prefix.
struct ProveScanTop {             // thread = proof
    static constexpr const char* kName = "ProveScanTop";
    PShape sh; PBuffers pb;
    CPG_HD void operator()(uint64_t b) const {
        HFr run = cpgh::fr_one();
        for (uint32_t k = 0; k < sh.nch; k++) { HFr* p = pb.part + (size_t)k * pb.B + b; HFr v = *p; *p = run; run = cpgh::fr_mul(run, v); }
    }
};
struct ProveScanApply {           // thread = (chunk, proof)
    static constexpr const char* kName = "ProveScanApply";
    PShape sh; PBuffers pb;
    CPG_HD void operator()(uint64_t t) const {
        const size_t b = (size_t)(t % pb.B); const uint32_t k = (uint32_t)(t / pb.B);
        FrVec fact = pvec(sh, pb, b, PV_FACT), c = pvec(sh, pb, b, PV_C);
        HFr run = pb.part[(size_t)k * pb.B + b];
        for (uint32_t i = k * sh.CH; i < (k + 1) * sh.CH && i < sh.ell; i++) { c[i] = run; run = cpgh::fr_mul(run, fact[i]); }
    }
};

// T_j = k R_perm[j], U_j = k S_perm[j]: thread = (proof, j in [0, 2 ell))            (curdleproofs.py:310-314)
// (One point per thread: a variant computing T_j and U_j in one thread to share their inversion ran 120 ms
// instead of 89.5 - two inlined scalar multiplications per kernel body miss the instruction cache.)
struct ProveShuffle {
    static constexpr const char* kName = "ProveShuffle";
    uint32_t ell;
    Aff* bases;                   // [B][4 ell]  R | S | T | U  (T, U written here)
    const uint32_t* perm;         // [B][ell]
    const uint32_t* k;            // [B][8]  k1 | k2 with k = k1 + k2 lambda (glv_split)
    uint8_t* tu48;                // [B][2 ell][48]
    CPG_HD void operator()(uint64_t t) const {
        uint64_t b = t / (2 * ell); uint32_t j = (uint32_t)(t % (2 * ell));
        Aff* row = bases + b * 4 * (uint64_t)ell;
        const uint32_t* pm = perm + b * (uint64_t)ell;
        Aff src = j < ell ? row[pm[j]] : row[ell + pm[j - ell]];
        Aff out = jac_to_aff(jac_mul_glv(to_jac(src), k + 8 * b));
        row[2 * (uint64_t)ell + j] = out;
        aff_compress(out, tu48 + 48 * t);
    }
};

// out[b][id] = compress(fixed[o][b] + scale * var[row][b]), thread = (output of the round, proof)
struct ProveCombine {
    static constexpr const char* kName = "ProveCombine";
    uint32_t nout, NOUT; uint64_t B;
    uint32_t out_id[P_MAX_OUT];
    int32_t var_row[P_MAX_OUT];   // -1: no variable-base part
    const uint32_t* scale[P_MAX_OUT];   // optional per-proof scalar multiplying the variable-base part ...
    uint32_t scale_stride[P_MAX_OUT];   // ... at scale[o] + b*scale_stride[o] (u32 words)
    const Jac* fixed;             // [nout][B]
    const Jac* var;               // [nvar][B]
    uint8_t* outs48;              // [B][NOUT][48]
    CPG_HD void operator()(uint64_t t) const {
        uint32_t o = (uint32_t)(t / B); uint64_t b = t % B;
        Jac p = fixed[t];
        if (var_row[o] >= 0) {
            Jac v = var[(uint64_t)var_row[o] * B + b];
            if (scale[o]) v = jac_mul(v, scale[o] + b * scale_stride[o]);
            p = jac_add(p, v);
        }
        aff_compress(jac_to_aff(p), outs48 + (b * NOUT + out_id[o]) * 48);
    }
};

// base offset (in points) of variable-base instance t = v*B + b: proof b's row [R|S|T|U], vector set[v]
// (tables = 1: offset, in bases, into the lane's [B][2 ell] table block T | U instead)
struct VarOffsets {
    static constexpr const char* kName = "VarOffsets";
    uint64_t B; uint32_t ell; uint32_t set[P_MAX_VAR]; uint32_t* off; int tables;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t v = t / B, b = t % B;
        off[t] = tables ? (uint32_t)(b * 2 * ell + (set[v] - 2) * ell) : (uint32_t)(b * 4 * ell + set[v] * ell);
    }
};

// One lane = the device buffers of a contiguous sub-batch.  A batch is split over `nlanes` lanes whose
// rounds are issued alternately on separate streams, so the latency-bound per-proof kernels of one lane
// (ProveTranscript: one thread per proof, 32 warps for 4096 proofs) run under the MSM kernels of the other.
// k = k1 + k2 lambda as integers, lambda = 0xac45a4010001a40200000000ffffffff (g1.cuh::jac_mul_glv); out = k1[4] | k2[4]
// returns false (and splits 0) when k is not a canonical scalar (k >= r): k2 would not fit 128 bits
bool glv_split(const uint8_t* k32, uint32_t* out) {
    uint64_t n[4];
    memcpy(n, k32, 32);                                  // little-endian host
    static const uint64_t R[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    bool lt = false;
    for (int i = 3; i >= 0; i--) { if (n[i] != R[i]) { lt = n[i] < R[i]; break; } }
    if (!lt) { memset(out, 0, 32); return false; }
    const unsigned __int128 lam = ((unsigned __int128)0xac45a4010001a402ULL << 64) | 0x00000000ffffffffULL;
    unsigned __int128 rem = 0, q = 0;
    for (int i = 255; i >= 0; i--) {
        const bool top = (rem >> 127) != 0;              // 2 rem + bit may pass 2^128: the subtraction below is exact mod 2^128
        rem = (rem << 1) | ((n[i >> 6] >> (i & 63)) & 1);
        const bool ge = top || rem >= lam;
        if (ge) rem -= lam;
        q = (q << 1) | (ge ? 1 : 0);
    }
    for (int i = 0; i < 4; i++) { out[i] = (uint32_t)(rem >> (32 * i)); out[4 + i] = (uint32_t)(q >> (32 * i)); }
    return true;
}

constexpr size_t TAB_CHUNK = 148 * 3 * 128 * 2;   // bases per VarTableBuild launch (bounds the scratch): two full waves of 3 blocks x 148 SMs
struct ProverLane {
    size_t cap = 0, B = 0;
    SideLane side;                // a few proofs: the round's fixed-base MSMs beside its variable-base ones (verify.inl)
    uint8_t *d_in48 = nullptr, *d_tu48 = nullptr, *d_k = nullptr, *d_rand = nullptr, *d_outs = nullptr, *d_fs = nullptr, *d_vs = nullptr, *d_proof = nullptr, *d_err = nullptr;
    uint32_t *d_perm = nullptr, *d_off = nullptr; Aff* d_bases = nullptr; HFr* d_vec = nullptr; Jac *d_fix = nullptr, *d_var = nullptr;
    Jac* d_gather = nullptr;      // [world][B * (P_MAX_OUT + P_MAX_VAR)] partial sums of a sharded proof's round
    PTr* d_trs = nullptr; PChal* d_chal = nullptr; PRes* d_res = nullptr; HFr *d_achal = nullptr, *d_part = nullptr, *d_red = nullptr;
    // host mirrors for the transcript on host threads (few, large proofs): pinned, exchanged once per round
    PChal* h_chal = nullptr; PRes* h_res = nullptr; HFr* h_achal = nullptr; std::vector<PTr> h_trs; std::vector<uint8_t> h_k;
    bool host_transcript = false;
    uint8_t *h_tu = nullptr, *h_outs = nullptr, *h_proof = nullptr, *h_err = nullptr;   // pinned: results leave per lane
    uint8_t *h_in = nullptr, *h_rand = nullptr;   // pinned staging of the two large inputs (pageable caller memory copies at a third of the rate)
    uint32_t* d_k12 = nullptr;    // [B][8] GLV halves of k
    Aff* d_tab = nullptr; uint32_t tab_ts = 0;   // multiples 1..tab_ts of every T_i, U_i: [B][2 ell][tab_s][tab_ts]
    uint32_t tab_s = 1;           // shift groups per base (VarTableBuild): 4 for a call of <= 16 proofs, whose rounds are latency chains
    Jac* d_tab_jac = nullptr; Fq* d_tab_pz = nullptr;   // build scratch for TAB_CHUNK bases at a time
#ifndef CPG_HOST_EMU
    cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
#endif
    std::vector<void*> all() { return {d_in48, d_tu48, d_k, d_rand, d_outs, d_fs, d_vs, d_proof, d_err, d_perm, d_off, d_bases, d_vec, d_fix, d_var, d_tab, d_tab_jac, d_tab_pz, d_k12,
                                        d_trs, d_chal, d_res, d_achal, d_part, d_red, d_gather}; }
    void release() {
        for (void* q : all()) cpg_free(q);
        cpg_host_free(h_tu); cpg_host_free(h_outs); cpg_host_free(h_proof); cpg_host_free(h_err); cpg_host_free(h_in); cpg_host_free(h_rand);
        cpg_host_free(h_chal); cpg_host_free(h_res); cpg_host_free(h_achal);
        h_chal = nullptr; h_res = nullptr; h_achal = nullptr; d_trs = nullptr; d_chal = nullptr; d_res = nullptr; d_achal = d_part = d_red = nullptr; d_gather = nullptr;
        h_tu = h_outs = h_proof = h_err = h_in = h_rand = nullptr;
        d_in48 = d_tu48 = d_k = d_rand = d_outs = d_fs = d_vs = d_proof = d_err = nullptr; d_perm = d_off = nullptr; d_bases = nullptr; d_vec = nullptr; d_fix = d_var = nullptr; d_tab = nullptr; d_tab_jac = nullptr; d_tab_pz = nullptr; d_k12 = nullptr;
        cap = 0;
    }
    int reserve(const PShape& sh, size_t proof_len, size_t Bn, uint32_t NOUT, uint32_t ts, int world = 1) {
        if (Bn <= cap && ts == tab_ts) return 0;
        release();
        tab_ts = ts;
        d_gather = (Jac*)cpg_malloc(sizeof(Jac) * (world > 1 ? (size_t)world * Bn * (P_MAX_OUT + P_MAX_VAR) : 1));
        d_k12 = (uint32_t*)cpg_malloc(Bn * 32);
        d_tab = (Aff*)cpg_malloc(sizeof(Aff) * (std::max<size_t>(Bn, 64) * 2 * sh.ell * (size_t)ts + 1));   // (>= 16 proofs x 4 shift groups)
        d_tab_jac = (Jac*)cpg_malloc(sizeof(Jac) * (TAB_CHUNK * (size_t)ts + 1));
        d_tab_pz = (Fq*)cpg_malloc(sizeof(Fq) * (TAB_CHUNK * (size_t)ts + 1));
        const size_t ell = sh.ell, n = sh.n;
        d_in48 = (uint8_t*)cpg_malloc(Bn * 2 * ell * 48);     d_tu48 = (uint8_t*)cpg_malloc(Bn * 2 * ell * 48);
        d_k = (uint8_t*)cpg_malloc(Bn * 32);                  d_rand = (uint8_t*)cpg_malloc(Bn * sh.NR * 32);
        d_outs = (uint8_t*)cpg_malloc(Bn * NOUT * 48);        d_fs = (uint8_t*)cpg_malloc(Bn * P_MAX_OUT * sh.NF * 32);
        d_vs = (uint8_t*)cpg_malloc(Bn * P_MAX_VAR * ell * 32); d_proof = (uint8_t*)cpg_malloc(Bn * proof_len);
        d_err = (uint8_t*)cpg_malloc(Bn * 2 * ell);           d_perm = (uint32_t*)cpg_malloc(Bn * ell * 4);
        d_off = (uint32_t*)cpg_malloc(Bn * P_MAX_VAR * 4);
        d_bases = (Aff*)cpg_malloc(sizeof(Aff) * Bn * 4 * ell);
        d_trs = (PTr*)cpg_malloc(sizeof(PTr) * Bn);           d_chal = (PChal*)cpg_malloc(sizeof(PChal) * Bn);
        d_res = (PRes*)cpg_malloc(sizeof(PRes) * Bn);         d_achal = (HFr*)cpg_malloc(sizeof(HFr) * Bn * ell);
        d_part = (HFr*)cpg_malloc(sizeof(HFr) * Bn * 2 * sh.nch); d_red = (HFr*)cpg_malloc(sizeof(HFr) * Bn * 2);
        h_chal = (PChal*)cpg_host_alloc(sizeof(PChal) * Bn); h_res = (PRes*)cpg_host_alloc(sizeof(PRes) * Bn);
        h_achal = (HFr*)cpg_host_alloc(sizeof(HFr) * Bn * ell);
        if (h_chal) memset(h_chal, 0, sizeof(PChal) * Bn);
        h_trs.assign(Bn, PTr());
        d_vec = (HFr*)cpg_malloc(sizeof(HFr) * Bn * PV_COUNT * n);
        d_fix = (Jac*)cpg_malloc(sizeof(Jac) * Bn * P_MAX_OUT); d_var = (Jac*)cpg_malloc(sizeof(Jac) * Bn * P_MAX_VAR);
        h_tu = (uint8_t*)cpg_host_alloc(Bn * 2 * ell * 48); h_outs = (uint8_t*)cpg_host_alloc(Bn * NOUT * 48);
        h_proof = (uint8_t*)cpg_host_alloc(Bn * proof_len);   h_err = (uint8_t*)cpg_host_alloc(Bn * 2 * ell);
        h_in = (uint8_t*)cpg_host_alloc(Bn * 2 * ell * 48);   h_rand = (uint8_t*)cpg_host_alloc(Bn * sh.NR * 32);
        for (void* q : all()) if (!q) { release(); return fail("cpg_prove_batch: device allocation failed"); }
        if (!h_tu || !h_outs || !h_proof || !h_err || !h_in || !h_rand || !h_chal || !h_res || !h_achal) { release(); return fail("cpg_prove_batch: pinned host allocation failed"); }
        cap = Bn;
        return 0;
    }
};
constexpr int P_MAX_LANES = 8;

struct Prover {
    PShape sh;
    size_t proof_len;             // 1088 + 480 lg (without M)
    std::vector<uint8_t> crs48;
    Aff* d_crs = nullptr; uint8_t* d_crs48 = nullptr;
    void* table = nullptr;        // fixed-base table over vec_G | vec_H | H | G_t | G_u
    Shard shard;                  // world > 1: ONE proof's leaves split over ranks (cpg_prover_create_sharded)
    std::vector<void*> shard_tables;   // [world] (only this rank's entry unless the ranks are emulated): table of that rank's block of the CRS bases
    int var_window = 0;
    int table_window = 6;         // per-base tables of 2^(c-1) multiples for the T / U MSMs (0: bucket method for those too)
    int nlanes = 2, lastK = 1;
    int transcript_mode = 2;      // 0 host threads, 1 one GPU thread per proof, 2 by batch size (cpg_prover_set_transcript)
    int threads = 1;
    // A CPU core runs a proof's Keccak chain ~20x faster than a lone GPU thread, the GPU runs thousands of chains at
    // once: the host takes the transcript while a lane holds fewer proofs than ~16 per host thread.
    bool host_transcript_for(size_t B) const { return transcript_mode == 0 || (transcript_mode == 2 && B <= 16 * (size_t)threads); }
    size_t lane_min = 256;        // proofs per lane below which a batch is not split
    size_t lastB = 0;
    ProverLane lanes[P_MAX_LANES];
#ifndef CPG_HOST_EMU
    cudaEvent_t fork = nullptr;
#endif
    void release() { for (ProverLane& L : lanes) L.release(); }
    // contiguous split of B proofs; small batches stay on one lane (nothing to hide behind)
    int split(size_t B, size_t* first, size_t* count) const {
        int k = (B >= lane_min * (size_t)nlanes && !shard.on()) ? nlanes : 1;   // (a sharded proof's collectives stay on one stream)
        for (int i = 0; i < k; i++) { first[i] = B * i / k; count[i] = B * (i + 1) / k - first[i]; }
        return k;
    }
};

// The device side of one lane (inputs already resident): decode + shuffle, then 8 + 2 lg rounds.
int prove_lane_prologue(Prover& pr, ProverLane& p) {
    const PShape sh = pr.sh;
    const uint32_t ell = sh.ell;
    const size_t B = p.B;
    // decode R | S into the first half of each proof's base row, then the shuffle itself
    {
        Scratch sc;
        Aff* tmp = sc.get<Aff>(B * 2 * (size_t)ell);
        if (!tmp) return fail("cpg_prove_batch: scratch allocation failed");
        if (int rc = cpg_g1_decompress(p.d_in48, B * 2 * (size_t)ell, 0, tmp, p.d_err)) return rc;
        // rows are [R|S|T|U]: place R|S at the start of each row
#ifndef CPG_HOST_EMU
        CK(cudaMemcpy2DAsync(p.d_bases, sizeof(Aff) * 4 * (size_t)ell, tmp, sizeof(Aff) * 2 * (size_t)ell, sizeof(Aff) * 2 * (size_t)ell, B,
                             cudaMemcpyDeviceToDevice, cur()));
#else
        for (size_t b = 0; b < B; b++) memcpy(p.d_bases + b * 4 * (size_t)ell, tmp + b * 2 * (size_t)ell, sizeof(Aff) * 2 * (size_t)ell);
#endif
    }
    if (int rc = launch(ProveShuffle{ell, p.d_bases, p.d_perm, p.d_k12, p.d_tu48}, B * 2 * (size_t)ell)) return rc;
    if (p.tab_ts) {                                     // multiples of every T_i, U_i (they enter 8 small MSMs each)
        // a call of a few proofs: 4 shift groups per base, so that the Horner pass of the 8 table-MSM rounds is 11 windows
        // (60 doublings, 0.4 ms) instead of 43 (258 doublings, 1.5 ms) - for the price of a 4x table and 198 doublings here
        p.tab_s = B <= 16 ? 4 : 1;
        const uint32_t c = (uint32_t)pr.table_window, Wt = windows_for(c), Wp = (Wt + p.tab_s - 1) / p.tab_s;
        const size_t nthreads = B * 2 * (size_t)ell * p.tab_s;
        for (size_t t0 = 0; t0 < nthreads; t0 += TAB_CHUNK) {
            size_t cnt = nthreads - t0 < TAB_CHUNK ? nthreads - t0 : TAB_CHUNK;
            if (int rc = launch_occ(VarTableBuild{p.tab_ts, p.d_bases, 4 * (uint64_t)ell, 2 * (uint64_t)ell, 2 * (uint64_t)ell, t0, p.d_tab_jac, p.d_tab_pz, p.d_tab, p.tab_s, c * Wp}, cnt)) return rc;
        }
    }
    return 0;
}
int prove_lane_round(Prover& pr, ProverLane& p, uint32_t r) {
    const PShape sh = pr.sh;
    const POut O(sh.lg);
    const uint32_t ell = sh.ell, lg = sh.lg;
    const size_t B = p.B;
    PBuffers pb;
    pb.in48 = p.d_in48; pb.tu48 = p.d_tu48; pb.perm = p.d_perm; pb.kbytes = p.d_k; pb.rand = p.d_rand; pb.crs48 = pr.d_crs48;
    pb.trs = p.d_trs; pb.chal = p.d_chal; pb.res = p.d_res; pb.achal = p.d_achal; pb.part = p.d_part; pb.red = p.d_red;
    pb.vec = p.d_vec; pb.outs48 = p.d_outs; pb.fs = p.d_fs; pb.vs = p.d_vs; pb.proof = p.d_proof; pb.B = B;
    const PRounds RD(lg);

    // per round: (first output id, count) and which outputs carry a variable-base part over which vector
    struct RoundPlan { uint32_t nout; uint32_t ids[P_MAX_OUT]; int32_t var_row[P_MAX_OUT]; uint32_t nvar; uint32_t var_set[P_MAX_VAR]; int scale_kind[P_MAX_OUT]; };
    auto plan_for = [&](uint32_t r) {
        RoundPlan pl; memset(&pl, 0, sizeof pl);
        for (uint32_t i = 0; i < P_MAX_OUT; i++) pl.var_row[i] = -1;
        auto add = [&](uint32_t id, int set) { pl.ids[pl.nout] = id; if (set >= 0) { pl.var_row[pl.nout] = (int32_t)pl.nvar; pl.var_set[pl.nvar++] = (uint32_t)set; } pl.nout++; };
        if (r == 0) add(O.M, -1);
        else if (r == 1) add(O.A, -1);
        else if (r == 2) add(O.B, -1);
        else if (r == 3) add(O.C, -1);
        else if (r == 4) { add(O.D, -1); add(O.Bc, -1); add(O.Bd, -1); }
        else if (r < 5 + lg) { for (uint32_t k = 0; k < 4; k++) add(O.ipa0 + 4 * (r - 5) + k, -1); }
        else if (r == 5 + lg) {
            // R' = MSM(vec_R, a) and S' = MSM(vec_S, a) are the only MSMs; cm_T2/cm_A2 reuse R' scaled by k / r_k,
            // cm_U2/cm_B2 reuse S' (curdleproofs.py:97-102, same_scalar.py:43-44)
            add(O.Rp, 0); add(O.Sp, 1); add(O.T1, -1); add(O.T2, -1); add(O.U1, -1); add(O.U2, -1); add(O.A1, -1); add(O.A2, -1); add(O.B1, -1); add(O.B2, -1);
            pl.var_row[3] = 0; pl.scale_kind[3] = 1; pl.var_row[5] = 1; pl.scale_kind[5] = 1;      // T2 = k R' + ..., U2 = k S' + ...
            pl.var_row[7] = 0; pl.scale_kind[7] = 2; pl.var_row[9] = 1; pl.scale_kind[9] = 2;      // A2 = r_k R' + ..., B2 = r_k S' + ...
        }
        else if (r == 6 + lg) { add(O.Ap, -1); add(O.Ba, -1); add(O.Bt, 2); add(O.Bu, 3); }
        else if (r < 7 + 2 * lg) { uint32_t base = O.msm0 + 6 * (r - 7 - lg); add(base, -1); add(base + 1, 2); add(base + 2, 3); add(base + 3, -1); add(base + 4, 2); add(base + 5, 3); }
        return pl;
    };
    // ---- transcript step: absorb the previous round's outputs, squeeze this round's challenges ----
    if (r > 0) {
        if (p.host_transcript) {
            // the previous round's compressed outputs and scalars come to the host, the challenges go back
            if (int rc = d2h_async(p.h_outs, p.d_outs, B * (size_t)O.NOUT * 48)) return rc;
            if (int rc = d2h_async(p.h_res, p.d_res, B * sizeof(PRes))) return rc;
            if (r == 1) if (int rc = d2h_async(p.h_tu, p.d_tu48, B * 2 * (size_t)ell * 48)) return rc;
            if (int rc = cpg_sync()) return rc;
            PTBuf pt;
            pt.in48 = p.h_in; pt.tu48 = p.h_tu; pt.kbytes = p.h_k.data(); pt.rand = p.h_rand; pt.crs48 = pr.crs48.data(); pt.outs48 = p.h_outs;
            pt.trs = p.h_trs.data(); pt.chal = p.h_chal; pt.res = p.h_res; pt.achal = p.h_achal; pt.proof = p.h_proof;
            parallel_for(pr.threads, B, [&](size_t b) { prove_transcript(sh, O, pt, r, b); });
            if (int rc = cpg_h2d(p.d_chal, p.h_chal, B * sizeof(PChal))) return rc;
            if (r == 1) if (int rc = cpg_h2d(p.d_achal, p.h_achal, B * (size_t)ell * sizeof(HFr))) return rc;
        } else {
            PTBuf pt;
            pt.in48 = p.d_in48; pt.tu48 = p.d_tu48; pt.kbytes = p.d_k; pt.rand = p.d_rand; pt.crs48 = pr.d_crs48; pt.outs48 = p.d_outs;
            pt.trs = p.d_trs; pt.chal = p.d_chal; pt.res = p.d_res; pt.achal = p.d_achal; pt.proof = p.d_proof;
            if (int rc = launch<64>(ProveTranscript{sh, O, pt, r}, B)) return rc;
            if (getenv("CPG_DEBUG_SHADOW")) {               // debugging aid: re-run the step on the host, compare the transcript states (this is how a
                                                            // device-only miscompilation of the identity's encoding was found; tools/prove_consistency.py)
                std::vector<PChal> dev(B);
                if (int rc = d2h_async(p.h_outs, p.d_outs, B * (size_t)O.NOUT * 48)) return rc;
                if (int rc = d2h_async(p.h_res, p.d_res, B * sizeof(PRes))) return rc;
                if (int rc = d2h_async(p.h_tu, p.d_tu48, B * 2 * (size_t)ell * 48)) return rc;
                if (int rc = cpg_d2h(dev.data(), p.d_chal, B * sizeof(PChal))) return rc;
                PTBuf ht;
                ht.in48 = p.h_in; ht.tu48 = p.h_tu; ht.kbytes = p.h_k.data(); ht.rand = p.h_rand; ht.crs48 = pr.crs48.data(); ht.outs48 = p.h_outs;
                ht.trs = p.h_trs.data(); ht.chal = p.h_chal; ht.res = p.h_res; ht.achal = p.h_achal; ht.proof = p.h_proof;
                for (size_t b = 0; b < B; b++) prove_transcript(sh, O, ht, r, b);
                size_t bad = 0;
                for (size_t b = 0; b < B; b++) if (memcmp(&dev[b], &p.h_chal[b], sizeof(PChal))) bad++;
                std::vector<PTr> dtr(B);
                if (int rc = cpg_d2h(dtr.data(), p.d_trs, B * sizeof(PTr))) return rc;
                size_t badtr = 0;
                for (size_t b = 0; b < B; b++) if (memcmp(dtr[b].tr.s.st.b, p.h_trs[b].tr.s.st.b, 200) || dtr[b].tr.s.pos != p.h_trs[b].tr.s.pos) badtr++;
                fprintf(stderr, "[shadow] round %u: %zu of %zu challenge sets differ, %zu transcript states differ\n", r, bad, B, badtr);
            }
        }
    }
    if (r == RD.FIN) return 0;
    // ---- vector kernels: this round's Fr vector work and coefficient rows, one thread per (proof, element) ----
    {
        auto V = [&](uint32_t phase) { return launch(ProveVec{sh, pb, r, phase}, (uint64_t)sh.NF * B); };
        auto R1 = [&](uint32_t kind, uint32_t m) { return launch(ProveReduce1{sh, pb, kind, m}, (uint64_t)sh.nch * B); };
        auto R2 = [&](uint32_t kind) { return launch(ProveReduce2{sh, pb, kind}, B); };
        int rc = 0;
        if (r == 2) { rc = V(0); if (!rc) rc = R1(PR_PROD_FACT, 0); if (!rc) rc = R2(PR_PROD_FACT); }
        else if (r == 3) {
            rc = R1(PR_PROD_FACT, 0);
            if (!rc) rc = launch(ProveScanTop{sh, pb}, B);
            if (!rc) rc = launch(ProveScanApply{sh, pb}, (uint64_t)sh.nch * B);
            if (!rc) rc = V(0);
        }
        else if (r == 4) { rc = V(0); if (!rc) rc = R1(PR_IPA_BLIND, 0); if (!rc) rc = R2(PR_IPA_BLIND); if (!rc) rc = V(1); }
        else if (r >= RD.IPA0 && r < RD.SS) {
            rc = V(0);
            if (!rc) rc = R1(PR_IPA_LR, (sh.n >> (r - RD.IPA0)) / 2);
            if (!rc) rc = R2(PR_IPA_LR);
            if (!rc) rc = V(1);
        }
        else if (r >= RD.MSM0) { rc = V(0); if (!rc) rc = V(1); }
        else rc = V(0);
        if (rc) return rc;
    }
    {
        RoundPlan pl = plan_for(r);
        if (pr.shard.on()) {
            // ONE proof over several ranks: each rank sums its block of the CRS leaves (its own table) and of the
            // tracker leaves (bucket method over a sub-range of every coefficient row); one all-gather per round
            const Shard& sd = pr.shard;
            const size_t nfix = B * pl.nout, nv = B * pl.nvar, cnt = nfix + nv;
            if (pl.nvar) {
                VarOffsets vo; vo.B = B; vo.ell = ell; vo.off = p.d_off; vo.tables = 0;
                for (uint32_t v = 0; v < P_MAX_VAR; v++) vo.set[v] = v < pl.nvar ? pl.var_set[v] : 0;
                if (int rc = launch(vo, nv)) return rc;
            }
            for (int rk = sd.first(); rk < sd.last(); rk++) {
                Jac* slot = p.d_gather + (size_t)rk * cnt;
                size_t lo, hi;
                comm_block(sh.NF, rk, sd.world, &lo, &hi);
                if (int rc = msm_fixed_impl(pr.shard_tables[rk], p.d_fs, nfix, 0, slot, sh.NF, lo)) return rc;
                if (pl.nvar) {
                    comm_block(ell, rk, sd.world, &lo, &hi);
                    if (int rc = msm_batched_impl(p.d_bases + lo, 0, p.d_off, p.d_vs, nv, hi - lo, pr.var_window, slot + nfix, 0, 0, ell, lo)) return rc;
                }
            }
            if (int rc = shard_exchange(sd, p.d_gather, cnt)) return rc;
            if (int rc = launch_occ(SumRanks{(uint32_t)sd.world, cnt, nfix, p.d_gather, p.d_fix, p.d_var}, cnt)) return rc;
        } else {
        // fixed-base part of every output of the round: B*nout MSMs over the CRS table (rows are output-major); for a
        // few proofs (B * nout <= 64 MSMs: both parts are latency chains) beside the variable-base part, not before it
        auto fixed_part = [&]() { return cpg_g1_msm_fixed_batched(pr.table, p.d_fs, B * pl.nout, 0, p.d_fix); };
        auto variable_part = [&]() -> int {
        if (pl.nvar) {                                  // variable-base parts: ONE batched MSM over all B*nvar instances
            VarOffsets vo; vo.B = B; vo.ell = ell; vo.off = p.d_off;
            for (uint32_t v = 0; v < P_MAX_VAR; v++) vo.set[v] = v < pl.nvar ? pl.var_set[v] : 0;
            bool tu_only = p.tab_ts != 0;               // every variable part of the round is over T or U: table look-ups
            for (uint32_t v = 0; v < pl.nvar; v++) tu_only = tu_only && pl.var_set[v] >= 2;
            vo.tables = tu_only ? 1 : 0;
            if (int rc = launch(vo, B * pl.nvar)) return rc;
            if (tu_only) {
                const uint32_t c = (uint32_t)pr.table_window;
                Recode rc_ = make_recode(c);
                const uint64_t M = B * pl.nvar;
                Scratch sc;
                // a few MSMs (one proof per call): 124 sequential additions per (msm, window) thread are 1.3 ms per round -
                // the bases are cut into chunks and the chunk sums added by one or two more short launches
                const uint32_t Wp = (rc_.W + p.tab_s - 1) / p.tab_s;          // windows after folding the table's shift groups
                uint32_t nchunk = 1;                                            // a power of two: ~16 additions per thread
                if (M * Wp < 8192 && ell >= 32) { const uint32_t want = (ell * p.tab_s + 15) / 16; while (nchunk < want && nchunk < 64) nchunk *= 2; }
                Xyzz* partial = sc.get<Xyzz>(M * Wp * nchunk);
                if (!partial) return fail("cpg_prove_batch: scratch allocation failed");
#ifndef CPG_HOST_EMU
                if (g_prof_on && g_d_counters) if (int rc = launch(CountNonZero{(const uint32_t*)p.d_vs, g_d_counters + 1, ell, ell, 0}, M * ell)) return rc;
#endif
                if (int rc = launch<128, 3>(VarTableMsmWindow{ell, p.tab_ts, rc_.W, rc_, (uint32_t)M, p.d_tab, p.d_off, (const uint32_t*)p.d_vs, partial, nchunk, p.tab_s, Wp},
                                            nchunk > 1 ? M * Wp * nchunk : ((M + 31) / 32) * 32 * Wp)) return rc;
                for (uint32_t nc = nchunk; nc > 1;) {                           // chunk sums: a tree of fan-in <= 8 (groups never straddle a window)
                    const uint32_t per = nc >= 8 ? 8 : nc, nout = nc / per;
                    Xyzz* sums = sc.get<Xyzz>(M * Wp * nout);
                    if (!sums) return fail("cpg_prove_batch: scratch allocation failed");
                    if (int rc = launch_occ(SumPartialsRagged{Wp * nc, per, Wp * nout, partial, sums}, M * Wp * nout)) return rc;
                    partial = sums; nc = nout;
                }
                MsmShape hs; memset(&hs, 0, sizeof hs); hs.W = Wp; hs.c = c;
                if (int rc = launch_horner_xyzz(hs, partial, p.d_var, M)) return rc;
            } else if (int rc = cpg_g1_msm_batched_off(p.d_bases, p.d_off, p.d_vs, B * pl.nvar, ell, pr.var_window, p.d_var)) return rc;
        }
        return 0;
        };
        if (int rc = p.side.beside(B * pl.nout <= 64 && pl.nvar != 0, fixed_part, variable_part)) return rc;
        }
        ProveCombine pc;
        pc.nout = pl.nout; pc.NOUT = O.NOUT; pc.B = B; pc.fixed = p.d_fix; pc.var = p.d_var; pc.outs48 = p.d_outs;
        const PRand RO(sh.n);
        for (uint32_t i = 0; i < P_MAX_OUT; i++) {
            pc.out_id[i] = pl.ids[i]; pc.var_row[i] = pl.var_row[i];
            pc.scale[i] = nullptr; pc.scale_stride[i] = 0;
            if (pl.scale_kind[i] == 1) { pc.scale[i] = (const uint32_t*)p.d_k; pc.scale_stride[i] = 8; }
            if (pl.scale_kind[i] == 2) { pc.scale[i] = (const uint32_t*)p.d_rand + (size_t)RO.r_k * 8; pc.scale_stride[i] = sh.NR * 8; }
        }
        if (int rc = launch_occ(pc, B * pl.nout)) return rc;
    }
    return 0;
}
// All lanes, rounds issued alternately.  The lane streams fork from the caller's stream and join it again,
// so events recorded on the caller's stream (cpg_timer_*) bracket the whole batch.
int prove_device_all(Prover& p, int k, bool download = false) {
    const uint32_t rounds = 8 + 2 * p.sh.lg;
    const POut O(p.sh.lg);
#ifndef CPG_HOST_EMU
    cudaStream_t caller = cur();
    const bool had = t_stream_set; cudaStream_t prev = t_stream;
    if (k > 1) {
        CK(cudaEventRecord(p.fork, caller));
        for (int i = 0; i < k; i++) CK(cudaStreamWaitEvent(p.lanes[i].stream, p.fork, 0));
    }
    auto enter = [&](int i) { if (k > 1) { t_stream = p.lanes[i].stream; t_stream_set = true; } };
    auto leave = [&]() { t_stream = prev; t_stream_set = had; };
#else
    auto enter = [&](int) {};
    auto leave = [&]() {};
#endif
    int rc = 0;
    for (int i = 0; i < k && !rc; i++) { enter(i); rc = prove_lane_prologue(p, p.lanes[i]); }
    for (uint32_t r = 0; r < rounds && !rc; r++)
        for (int i = 0; i < k && !rc; i++) { enter(i); rc = prove_lane_round(p, p.lanes[i], r); }
    if (download)                                           // each lane's results leave on its own stream, under the other lanes' last rounds
        for (int i = 0; i < k && !rc; i++) {
            ProverLane& L = p.lanes[i];
            enter(i);
            rc = d2h_async(L.h_tu, L.d_tu48, L.B * 2 * (size_t)p.sh.ell * 48);
            if (!rc) rc = d2h_async(L.h_outs, L.d_outs, L.B * (size_t)O.NOUT * 48);
            if (!rc && !L.host_transcript) rc = d2h_async(L.h_proof, L.d_proof, L.B * p.proof_len);   // (host transcript: assembled in place)
            if (!rc) rc = d2h_async(L.h_err, L.d_err, L.B * 2 * (size_t)p.sh.ell);
        }
    leave();
#ifndef CPG_HOST_EMU
    if (k > 1) for (int i = 0; i < k; i++) {
        CK(cudaEventRecord(p.lanes[i].done, p.lanes[i].stream));
        CK(cudaStreamWaitEvent(caller, p.lanes[i].done, 0));
    }
#endif
    return rc;
}

}  // namespace

extern "C" {

static void* prover_create_impl(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, bool sharded);
void* cpg_prover_create(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window) {
    return prover_create_impl(crs_bytes, ell, n_blinders, fixed_window, false);
}
/* ONE proof at a time over ALL ranks of the communicator (BASELINE config 5): every rank calls this and every later
 * cpg_prove_batch with identical arguments and gets identical outputs; the proof's leaves are split over the ranks
 * (this rank builds the CRS table of its block only) and each round's partial sums cross by one all-gather. */
void* cpg_prover_create_sharded(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window) {
    return prover_create_impl(crs_bytes, ell, n_blinders, fixed_window, true);
}
static void* prover_create_impl(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, bool sharded) {
    if (need_init()) return nullptr;
    size_t n = ell + n_blinders;
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    if (((size_t)1 << lg) != n || n_blinders != 4 || ell < 4 || lg > MAX_LG) { fail("cpg_prover_create: need ell + 4 = 2^k, 3 <= k <= 16"); return nullptr; }
    Prover* p = new Prover;
    p->sh.ell = (uint32_t)ell; p->sh.n = (uint32_t)n; p->sh.lg = lg; p->sh.NF = (uint32_t)n + 3;
    p->sh.NR = PRand((uint32_t)n).NR; p->sh.rounds = 8 + 2 * lg;
    p->sh.CH = 16; while ((size_t)p->sh.CH * p->sh.CH < n) p->sh.CH *= 2;      // ~sqrt(n): both passes of a reduction stay short
    p->sh.nch = (uint32_t)((n + p->sh.CH - 1) / p->sh.CH);
    p->threads = (int)std::max(1u, std::thread::hardware_concurrency());
    p->proof_len = 1088 + 480 * (size_t)lg;
    p->crs48.assign(crs_bytes, crs_bytes + 48 * (n + 5));
    p->d_crs48 = (uint8_t*)cpg_malloc(48 * (n + 5));
    uint8_t* derr = (uint8_t*)cpg_malloc(n + 5);
    p->d_crs = (Aff*)cpg_malloc(sizeof(Aff) * (n + 5));
    std::vector<uint8_t> err(n + 5, 1);
    int rc = (!p->d_crs48 || !derr || !p->d_crs) ? fail("cpg_prover_create: allocation failed") : 0;
    if (!rc) rc = cpg_h2d(p->d_crs48, crs_bytes, 48 * (n + 5));
    if (!rc) rc = cpg_g1_decompress(p->d_crs48, n + 5, 0, p->d_crs, derr);
    if (!rc) rc = cpg_d2h(err.data(), derr, n + 5);
    cpg_free(derr);
    if (!rc) for (uint8_t e : err) if (e) { rc = fail("cpg_prover_create: CRS holds an invalid point encoding"); break; }
    if (sharded) p->shard = shard_now();
    if (!rc && p->shard.on()) {
        p->shard_tables.assign(p->shard.world, nullptr);
        for (int rk = p->shard.first(); rk < p->shard.last() && !rc; rk++) {
            size_t lo, hi;
            comm_block(n + 3, rk, p->shard.world, &lo, &hi);
            if (hi == lo) { rc = fail("cpg_prover_create_sharded: more ranks than CRS bases"); break; }
            p->shard_tables[rk] = cpg_fixed_table_create(p->d_crs + lo, hi - lo, fixed_window > 0 ? fixed_window : 12);
            if (!p->shard_tables[rk]) rc = 1;
        }
    } else if (!rc) { p->table = cpg_fixed_table_create(p->d_crs, n + 3, fixed_window > 0 ? fixed_window : 12); if (!p->table) rc = 1; }
#ifndef CPG_HOST_EMU
    if (!rc) rc = ck(cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming), "cudaEventCreate");
    for (ProverLane& L : p->lanes) {
        if (!rc) rc = ck(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = ck(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming), "cudaEventCreate");
    }
#endif
    if (rc) { cpg_prover_free(p); return nullptr; }
    return p;
}

int cpg_prover_free(void* handle) {
    if (!handle) return 0;
    Prover* p = (Prover*)handle;
    p->release();
#ifndef CPG_HOST_EMU
    for (ProverLane& L : p->lanes) { if (L.stream) cudaStreamDestroy(L.stream); if (L.done) cudaEventDestroy(L.done); L.side.destroy(); }
    if (p->fork) cudaEventDestroy(p->fork);
#endif
    cpg_fixed_table_free(p->table);
    for (void* t : p->shard_tables) cpg_fixed_table_free(t);
    cpg_free(p->d_crs); cpg_free(p->d_crs48);
    delete p;
    return 0;
}
size_t cpg_prover_proof_bytes(const void* handle) { return handle ? ((const Prover*)handle)->proof_len + 48 : 0; }
size_t cpg_prover_rand_scalars(const void* handle) { return handle ? ((const Prover*)handle)->sh.NR : 0; }
/* re-run the device side of the last cpg_prove_batch on its resident inputs (timing with inputs in HBM) */
int cpg_prove_replay_device(void* handle) {
    NEED_INIT();
    if (!handle) return fail("cpg_prove_replay_device: null prover");
    Prover& p = *(Prover*)handle;
    if (!p.lastB) return fail("cpg_prove_replay_device: no batch resident");
    return prove_device_all(p, p.lastK);
}
int cpg_prover_set_window(void* handle, int w) { if (!handle) return 1; ((Prover*)handle)->var_window = w; return 0; }
int cpg_prover_set_table_window(void* handle, int window) {
    if (!handle) return fail("cpg_prover_set_table_window: null prover");
    if (window < 0 || window > 10) return fail("cpg_prover_set_table_window: 0 (off) or 2..10");
    ((Prover*)handle)->table_window = window == 1 ? 0 : window;
    return 0;
}
int cpg_prover_set_transcript(void* handle, int mode) {
    if (!handle) return fail("cpg_prover_set_transcript: null prover");
    if (mode < 0 || mode > 2) return fail("cpg_prover_set_transcript: 0 (host threads), 1 (GPU thread per proof) or 2 (by batch size)");
    ((Prover*)handle)->transcript_mode = mode;
    return 0;
}
int cpg_prover_set_lanes(void* handle, int nlanes, size_t min_proofs_per_lane) {
    if (!handle) return fail("cpg_prover_set_lanes: null prover");
    if (nlanes < 1 || nlanes > P_MAX_LANES) return fail("cpg_prover_set_lanes: 1..8 lanes");
    ((Prover*)handle)->nlanes = nlanes;
    ((Prover*)handle)->lane_min = min_proofs_per_lane ? min_proofs_per_lane : 256;
    return 0;
}

/* inputs   : [B][2*ell*48]  vec_R | vec_S            (pre-shuffle tracker halves)
 * perms    : [B][ell] u32   permutation (post[j] = k * pre[perm[j]])
 * ks       : [B][32]        the shuffle scalar k
 * rand     : [B][NR][32]    blinders in the reference's draw order:
 *                           m_bl(4) a_bl(2) c_bl(4) ipa_r(n) ipa_z(n-2) r_t r_u r_a r_b r_k msm_r(n)
 * out_tu   : [B][2*ell*48]  vec_T | vec_U            (post-shuffle tracker halves)
 * out_proofs:[B][48 + 1088 + 480 lg]  M | proof      (WhiskShuffleProof.to_bytes)
 * status   : [B]  0 ok, 1 malformed input encoding (that lane's outputs are undefined) */
int cpg_prove_batch(void* handle, const uint8_t* inputs, const uint32_t* perms, const uint8_t* ks, const uint8_t* rand,
                    size_t B, uint8_t* out_tu, uint8_t* out_proofs, uint8_t* status) {
    NEED_INIT();
    if (!handle) return fail("cpg_prove_batch: null prover");
    if (!B) return 0;
    Prover& p = *(Prover*)handle;
    const PShape sh = p.sh;
    const POut O(sh.lg);
    const uint32_t ell = sh.ell;
    size_t first[P_MAX_LANES], count[P_MAX_LANES];
    const int k = p.split(B, first, count);
    const int threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<uint32_t> k12(B * 8);
    std::vector<uint8_t> bad_k(B, 0);
    // Caller-supplied indices and scalars are checked before anything reaches the device: perm entries index the base
    // rows and the challenge vector, so each row must be a true permutation of [0, ell) (the reference raises IndexError
    // otherwise), and k / every blinder must be a canonical scalar (Scalar.from_le_bytes raises ValueError).  A bad lane
    // is flagged in status[] and proves over a harmless substitute (identity permutation, blinder 1).
    std::vector<uint32_t> perm_ok(B * (size_t)ell);
    parallel_for(threads, B, [&](size_t b) {
        bad_k[b] = glv_split(ks + 32 * b, k12.data() + 8 * b) ? 0 : 1;
        const uint32_t* pm = perms + b * (size_t)ell;
        uint32_t* out = perm_ok.data() + b * (size_t)ell;
        std::vector<uint8_t> seen(ell, 0);
        bool ok = true;
        for (uint32_t j = 0; j < ell && ok; j++) { ok = pm[j] < ell && !seen[pm[j]]; if (ok) seen[pm[j]] = 1; }
        for (uint32_t j = 0; j < ell; j++) out[j] = ok ? pm[j] : j;
        if (!ok) bad_k[b] = 1;
    });
    for (int i = 0; i < k; i++) {
        ProverLane& L = p.lanes[i];
        const size_t f = first[i], c = count[i];
        // per-base tables pay off for the many small MSMs of Whisk-size proofs; one thread walks all ell terms of a
        // (msm, window), so large shuffles keep the bucket method
        if (int rc = L.reserve(sh, p.proof_len, c, O.NOUT, (p.table_window > 0 && ell <= 2048 && !p.shard.on()) ? 1u << (p.table_window - 1) : 0, p.shard.world)) return rc;
        L.B = c;
        L.host_transcript = p.host_transcript_for(c);
        L.h_k.assign(ks + f * 32, ks + (f + c) * 32);
        // the two large inputs go through pinned staging, copied by all host threads
        const size_t in_row = 2 * (size_t)ell * 48, rand_row = (size_t)sh.NR * 32;
        parallel_for(threads, c, [&](size_t j) {
            memcpy(L.h_in + j * in_row, inputs + (f + j) * in_row, in_row);
            uint8_t* rr = L.h_rand + j * rand_row;
            memcpy(rr, rand + (f + j) * rand_row, rand_row);
            for (uint32_t q = 0; q < sh.NR; q++) {
                uint64_t w[4];
                memcpy(w, rr + 32 * (size_t)q, 32);
                if (cpgh::fr_geq_mod(w)) { memset(rr + 32 * (size_t)q, 0, 32); rr[32 * (size_t)q] = 1; bad_k[f + j] = 1; }
            }
        });
        if (int rc = cpg_h2d(L.d_in48, L.h_in, c * in_row)) return rc;
        if (int rc = cpg_h2d(L.d_perm, perm_ok.data() + f * (size_t)ell, c * (size_t)ell * 4)) return rc;
        if (int rc = cpg_h2d(L.d_k, ks + f * 32, c * 32)) return rc;
        if (int rc = cpg_h2d(L.d_k12, k12.data() + f * 8, c * 32)) return rc;
        if (int rc = cpg_h2d(L.d_rand, L.h_rand, c * rand_row)) return rc;
    }
    p.lastB = B; p.lastK = k;
    if (int rc = prove_device_all(p, k, true)) return rc;
    if (int rc = cpg_sync()) return rc;
    // results: T|U, M|proof, per-lane status, straight from the lanes' pinned buffers
    for (int i = 0; i < k; i++) {
        ProverLane& L = p.lanes[i];
        const size_t f = first[i], c = count[i];
        parallel_for(threads, c, [&](size_t j) {
            const size_t b = f + j;
            memcpy(out_tu + b * 2 * (size_t)ell * 48, L.h_tu + j * 2 * (size_t)ell * 48, 2 * (size_t)ell * 48);
            uint8_t* w = out_proofs + b * (p.proof_len + 48);
            memcpy(w, L.h_outs + (j * O.NOUT + O.M) * 48, 48);
            memcpy(w + 48, L.h_proof + j * p.proof_len, p.proof_len);
            uint8_t bad = bad_k[b];
            const uint8_t* e = L.h_err + j * 2 * (size_t)ell;
            for (size_t q = 0; q < 2 * (size_t)ell; q++) bad |= e[q];
            if (status) status[b] = bad ? 1 : 0;
        });
    }
    return 0;
}

}  // extern "C"

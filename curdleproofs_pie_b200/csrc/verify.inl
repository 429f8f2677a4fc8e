// Batched shuffle-proof verification: B independent Whisk-size proofs per call.
// Included at the end of cpg_api.cu (same translation unit: uses launch(), cpg_* helpers).
//
// Replaces, for a whole batch at once, the reference's per-proof path
//   IsValidWhiskShuffleProof          curdleproofs/curdleproofs/whisk_interface.py:74-108
//   CurdleProofsProof.verify          curdleproofs/curdleproofs/curdleproofs.py:162-248
//   SamePermutation/GrandProduct/IPA/SameScalar/SameMSM .verify
//                                     same_perm.py:74-120, grand_prod.py:121-177, ipa.py:155-233,
//                                     same_scalar.py:71-111, same_msm.py:146-226
//   MSMAccumulator                    msm_accumulator.py:32-68
// Per proof the work is
//   phase 1  transcript up to gprod_beta                               (verify_phase1)
//   derive   D = B - beta^-1 G_sum + alpha H_sum,  A' = A + T_1 + U_1  (GPU: 2-base fixed MSM + VerifyDerived)
//   phase 2  rest of the transcript, then the MSM coefficients          (verify_phase2)
//   check    ONE MSM per proof: the reference's accumulator folds its 8 checks with random weights
//            (msm_accumulator.py:37-58); we fold the same 8 plus SameScalar's 4 equalities and expand
//            every left-hand side into proof points, so the verdict is
//            "MSM over (CRS | inputs | proof points | D) == identity".  CRS terms go through
//            fixed-base tables, the rest through the batched bucket method.
// Every group operation runs on the GPU.  The two transcript phases are ONE CPG_HD implementation
// (host_transcript.h) with two placements: host threads, one proof per task (transcript_on_device
// = 0, the north star's placement), or one proof per GPU thread (= 1, SURVEY 8 f-1: removes the
// host cores as the end-to-end bound).  Verdicts equal the reference's except with probability
// ~2^-250 over the batching weights (the reference's own accumulator has the same soundness error).
#include "host_transcript.h"

#include <functional>
#include <thread>
#if defined(__linux__)
#include <sys/random.h>
#endif

namespace {

using cpgh::HFr;

// offsets of the proof's points (index into the gathered point list) in the wire format
//   M | A | cm_T | cm_U | R | S | B | C | r_p | B_c | B_d | L_C R_C L_D R_D | c d |
//   cm_A | cm_B | z_k z_t z_u | B_a B_t B_u | L_A L_T L_U R_A R_T R_U | x          
struct Layout {
    uint32_t lg;
    uint32_t A, T1, T2, U1, U2, R, S, B, C, Bc, Bd;
    uint32_t LC, RC, LD, RD, A1, A2, B1, B2, Ba, Bt, Bu, LA, LT, LU, RA, RT, RU;
    explicit Layout(uint32_t lg_) : lg(lg_) {
        A = 0; T1 = 1; T2 = 2; U1 = 3; U2 = 4; R = 5; S = 6; B = 7; C = 8; Bc = 9; Bd = 10;
        LC = 11; RC = LC + lg; LD = RC + lg; RD = LD + lg;
        A1 = RD + lg; A2 = A1 + 1; B1 = A2 + 1; B2 = B1 + 1;
        Ba = B2 + 1; Bt = Ba + 1; Bu = Bt + 1;
        LA = Bu + 1; LT = LA + lg; LU = LT + lg; RA = LU + lg; RT = RA + lg; RU = RT + lg;
    }
};

constexpr uint32_t MAX_LG = 16;

struct VShape { uint32_t ell, n, lg, NP, NI, NV, NF; };

// per-proof state carried from phase 1 to phase 2
struct VState {
    cpgh::Transcript tr;
    HFr alpha_sp, beta_sp, gprod, alpha_gp, beta_gp, beta_gp_inv;
    HFr r_p, c_final, d_final, z_k, z_t, z_u, x_final;
    uint32_t bad;                 // malformed scalar encoding
};

// All pointers live in the memory space of whoever runs the phases (host vectors or device buffers).
struct VBuffers {
    const uint8_t* wire;          // [B][NV][48]  R|S|T|U|M|proof points|(D slot)
    const uint8_t* psc;           // [B][7][32]   proof scalars r_p c d z_k z_t z_u x
    const uint8_t* crs48;         // CRS wire bytes (H is appended to the transcript)
    VState* st;                   // [B]
    HFr* a;                       // [B][ell]     vec_a challenges
    HFr* tmp;                     // [B][5n]      s1 | s1^-1 | s2 | batch-inversion scratch (2n)
    uint8_t* chal;                // [B][64]      -beta^-1 | alpha  (scalars of the D MSM)
    const uint8_t* derived;       // [B][96]      compress(D) | compress(A')
    const uint8_t* err;           // [B][NV]      decode error codes
    const uint8_t* t0;            // [B]          vec_T[0] is the identity
    uint8_t* vs;                  // [B][NV][32]  variable-base coefficients (out)
    uint8_t* fs;                  // [B][NF][32]  fixed-base coefficients (out)
    uint8_t* reject;              // [B]          structural reject (out)
    uint8_t secret[32];
    uint64_t nonce;               // per-call counter: identical (proof, lane) pairs do not reuse weights across calls
};

CPG_HD void verify_phase1(const VShape& sh, const Layout& L, const VBuffers& vb, size_t b, uint32_t warp = 0) {
    using namespace cpgh;
    VState& s = vb.st[b];
    const uint32_t ell = sh.ell;
    const uint8_t* row = vb.wire + b * (size_t)sh.NV * 48;
    const uint8_t* pp = row + (size_t)sh.NI * 48;
    const uint8_t* sc = vb.psc + b * 7 * 32;
    HFr* a = vb.a + b * (size_t)ell;
    bool ok = fr_from_bytes(&s.r_p, sc) && fr_from_bytes(&s.c_final, sc + 32) && fr_from_bytes(&s.d_final, sc + 64) &&
              fr_from_bytes(&s.z_k, sc + 96) && fr_from_bytes(&s.z_t, sc + 128) && fr_from_bytes(&s.z_u, sc + 160) &&
              fr_from_bytes(&s.x_final, sc + 192);
    s.bad = ok ? 0 : 1;
    if (!ok) { memset(vb.chal + b * 64, 0, 64); return; }
    // the STROBE state is worked on in thread-local storage (registers / coalesced local memory) and
    // written back once: per-thread structs in global memory make every byte XOR an uncoalesced RMW
    Transcript tr;
    tr.s.warp = (uint8_t)warp;
    tr.init("curdleproofs");
    for (uint32_t i = 0; i < 4 * ell; i++) tr.append_point("curdleproofs_step1", row + 48 * (size_t)i);
    const uint8_t* M = row + 48 * (size_t)(4 * ell);
    tr.append_point("curdleproofs_step1", M);
    for (uint32_t i = 0; i < ell; i++) a[i] = tr.challenge("curdleproofs_vec_a");
    tr.append_point("same_perm_step1", pp + 48 * L.A);
    tr.append_point("same_perm_step1", M);
    for (uint32_t i = 0; i < ell; i++) tr.append_fr("same_perm_step1", a[i]);
    s.alpha_sp = tr.challenge("same_perm_alpha");
    s.beta_sp = tr.challenge("same_perm_beta");
    HFr g = fr_one(), ia = fr_zero();                                    // ia = i * alpha
    for (uint32_t i = 0; i < ell; i++) {
        g = fr_mul(g, fr_add(fr_add(a[i], ia), s.beta_sp));
        ia = fr_add(ia, s.alpha_sp);
    }
    s.gprod = g;
    tr.append_point("gprod_step1", pp + 48 * L.B);
    tr.append_fr("gprod_step1", s.gprod);
    s.alpha_gp = tr.challenge("gprod_alpha");
    tr.append_point("gprod_step2", pp + 48 * L.C);
    tr.append_fr("gprod_step2", s.r_p);
    s.beta_gp = tr.challenge("gprod_beta");
    s.beta_gp_inv = fr_inv(s.beta_gp);
    fr_to_bytes(vb.chal + b * 64, fr_neg(s.beta_gp_inv));
    fr_to_bytes(vb.chal + b * 64 + 32, s.alpha_gp);
    s.tr = tr;
}

CPG_HD void verify_phase2(const VShape& sh, const Layout& L, const VBuffers& vb, size_t b, uint32_t warp = 0) {
    using namespace cpgh;
    VState& s = vb.st[b];
    const uint32_t ell = sh.ell, n = sh.n, lg = sh.lg, NV = sh.NV, NF = sh.NF, NI = sh.NI;
    uint8_t* vrow = vb.vs + b * (size_t)NV * 32;
    uint8_t* frow = vb.fs + b * (size_t)NF * 32;
    zero_bytes(vrow, (size_t)NV * 32);
    zero_bytes(frow, (size_t)NF * 32);
    bool rej = s.bad || vb.t0[b];
    const uint8_t* e = vb.err + b * (size_t)NV;
    for (uint32_t i = 0; i + 1 < NV && !rej; i++) if (e[i]) rej = true;   // any malformed point encoding
    vb.reject[b] = rej ? 1 : 0;
    if (rej) return;
    const HFr* a = vb.a + b * (size_t)ell;
    HFr* s1 = vb.tmp + b * (size_t)(5 * n);
    HFr* s1i = s1 + n;
    HFr* s2 = s1i + n;
    HFr* scratch = s2 + n;
    const uint8_t* row = vb.wire + b * (size_t)NV * 48;
    const uint8_t* pp = row + (size_t)NI * 48;
    const uint8_t* Dbytes = vb.derived + b * 96;
    const uint8_t* Apbytes = Dbytes + 48;
    // (a constant, not a local array: nvcc 12.9 miscompiled `uint8_t INF[48]; memset; INF[0] = 0xc0` in the prover's
    // transcript step for sm_100a - the appended bytes were not the identity's - while the same source was right on the host)
    const uint8_t* INF = CPGH_SEL(INF48);
    Transcript tr = s.tr;
    tr.s.warp = (uint8_t)warp;
    // grand product -> IPA statement
    HFr beta_l = fr_pow_u64(s.beta_gp, ell), beta_l1 = fr_mul(beta_l, s.beta_gp);
    HFr z = fr_sub(fr_add(fr_mul(s.r_p, beta_l1), fr_mul(s.gprod, beta_l)), fr_one());
    tr.append_point("ipa_step1", pp + 48 * L.C);
    tr.append_point("ipa_step1", Dbytes);
    tr.append_fr("ipa_step1", z);
    tr.append_point("ipa_step1", pp + 48 * L.Bc);
    tr.append_point("ipa_step1", pp + 48 * L.Bd);
    HFr alpha_ipa = tr.challenge("ipa_alpha"), beta_ipa = tr.challenge("ipa_beta");
    HFr gam[MAX_LG], gam_inv[MAX_LG], gam2[MAX_LG], gam2_inv[MAX_LG];
    for (uint32_t j = 0; j < lg; j++) {
        tr.append_point("ipa_loop", pp + 48 * (L.LC + j));
        tr.append_point("ipa_loop", pp + 48 * (L.LD + j));
        tr.append_point("ipa_loop", pp + 48 * (L.RC + j));
        tr.append_point("ipa_loop", pp + 48 * (L.RD + j));
        gam[j] = tr.challenge("ipa_gamma");
    }
    // same scalar
    const uint32_t ss[10] = {L.R, L.S, L.T1, L.T2, L.U1, L.U2, L.A1, L.A2, L.B1, L.B2};
    for (uint32_t k = 0; k < 10; k++) tr.append_point("sameexp_points", pp + 48 * ss[k]);
    HFr alpha_ss = tr.challenge("same_scalar_alpha");
    // same MSM
    tr.append_point("same_msm_step1", Apbytes);
    tr.append_point("same_msm_step1", pp + 48 * L.T2);
    tr.append_point("same_msm_step1", pp + 48 * L.U2);
    const uint8_t* Hb = vb.crs48 + 48 * (size_t)n;                       // H follows vec_G | vec_H
    const uint8_t* Tb = row + 48 * (size_t)(2 * ell);
    const uint8_t* Ub = row + 48 * (size_t)(3 * ell);
    for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", Tb + 48 * (size_t)i);
    tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
    tr.append_point("same_msm_step1", Hb); tr.append_point("same_msm_step1", INF);
    for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", Ub + 48 * (size_t)i);
    tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
    tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", Hb);
    tr.append_point("same_msm_step1", pp + 48 * L.Ba);
    tr.append_point("same_msm_step1", pp + 48 * L.Bt);
    tr.append_point("same_msm_step1", pp + 48 * L.Bu);
    HFr alpha_msm = tr.challenge("same_msm_alpha");
    for (uint32_t j = 0; j < lg; j++) {
        tr.append_point("same_msm_loop", pp + 48 * (L.LA + j));
        tr.append_point("same_msm_loop", pp + 48 * (L.LT + j));
        tr.append_point("same_msm_loop", pp + 48 * (L.LU + j));
        tr.append_point("same_msm_loop", pp + 48 * (L.RA + j));
        tr.append_point("same_msm_loop", pp + 48 * (L.RT + j));
        tr.append_point("same_msm_loop", pp + 48 * (L.RU + j));
        gam2[j] = tr.challenge("same_msm_gamma");
    }
    // batching weights: 8 accumulator checks (rho) + 4 SameScalar equalities (delta), bound to the
    // whole transcript, the lane and a per-process secret so a prover cannot predict them
    HFr w[12];
    {
        Transcript fork = tr;
        fork.append("cpg_batch_secret", vb.secret, 32);
        uint64_t lane = (uint64_t)b;
        uint8_t lb[16];
        for (int k = 0; k < 8; k++) { lb[k] = (uint8_t)(lane >> (8 * k)); lb[8 + k] = (uint8_t)(vb.nonce >> (8 * k)); }
        fork.append("cpg_batch_lane", lb, 16);
        uint8_t raw[12 * 32];
        fork.challenge_bytes("cpg_batch_weights", raw, sizeof raw);
        for (int k = 0; k < 12; k++) {
            raw[32 * k + 31] &= 0x3f;                                     // < 2^254 < r
            if (!fr_from_bytes(&w[k], raw + 32 * k) || fr_is_zero(w[k])) w[k] = fr_one();
        }
    }
    const HFr &rho1 = w[0], &rho2 = w[1], &rho3 = w[2], &rho5 = w[3], &rho6 = w[4], &rho7 = w[5], &rho8 = w[6], &rho9 = w[7];
    const HFr &dl1 = w[8], &dl2 = w[9], &dl3 = w[10], &dl4 = w[11];
    // inverses of all round challenges in one batch
    {
        HFr inv[2 * MAX_LG];
        for (uint32_t j = 0; j < lg; j++) { inv[j] = gam[j]; inv[lg + j] = gam2[j]; }
        fr_batch_inv(inv, 2 * (size_t)lg, scratch);
        for (uint32_t j = 0; j < lg; j++) { gam_inv[j] = inv[j]; gam2_inv[j] = inv[lg + j]; }
    }
    // s-vectors: s_i = prod_{j: bit (lg-1-j) of i set} gamma_j  (util.py:71-78); s_i^-1 likewise
    s1[0] = s1i[0] = s2[0] = fr_one();
    for (uint32_t j = 0; j < lg; j++) {
        uint32_t half = 1u << j;
        const HFr &g1 = gam[lg - 1 - j], &g1i = gam_inv[lg - 1 - j], &g2 = gam2[lg - 1 - j];
        for (uint32_t i = 0; i < half; i++) {
            s1[half + i] = fr_mul(s1[i], g1);
            s1i[half + i] = fr_mul(s1i[i], g1i);
            s2[half + i] = fr_mul(s2[i], g2);
        }
    }
    // ---- fixed (CRS) coefficients: vec_G | vec_H | H | G_t | G_u ----
    HFr c2 = fr_mul(rho2, s.c_final), c3 = fr_mul(rho3, s.d_final), c5 = fr_mul(rho5, s.x_final);
    HFr c6 = fr_mul(rho6, s.x_final), c7 = fr_mul(rho7, s.x_final);
    HFr r1b = fr_mul(rho1, s.beta_sp);
    HFr u = s.beta_gp_inv;                                               // u_i = beta^-(i+1)
    HFr u_bl = fr_pow_u64(s.beta_gp_inv, ell + 1);
    for (uint32_t i = 0; i < n; i++) {
        const HFr& ui = i < ell ? u : u_bl;
        HFr t = fr_add(fr_mul(c2, s1[i]), fr_mul(c3, fr_mul(s1i[i], ui)));
        if (i < ell) t = fr_add(t, r1b);
        if (i < ell + 2) t = fr_add(t, fr_mul(c5, s2[i]));                 // G_wb = vec_G | vec_H[:2] | G_t | G_u
        fr_to_bytes(frow + 32 * (size_t)i, fr_neg(t));
        if (i < ell) u = fr_mul(u, s.beta_gp_inv);
    }
    {
        HFr t = fr_mul(rho2, fr_mul(beta_ipa, fr_sub(fr_mul(fr_mul(alpha_ipa, alpha_ipa), z), fr_mul(s.c_final, s.d_final))));
        t = fr_add(t, fr_add(fr_mul(dl2, s.z_t), fr_mul(dl4, s.z_u)));
        t = fr_sub(t, fr_add(fr_mul(c6, s2[ell + 2]), fr_mul(c7, s2[ell + 3])));
        fr_to_bytes(frow + 32 * (size_t)n, t);                                                            // H
        fr_to_bytes(frow + 32 * (size_t)(n + 1), fr_sub(fr_mul(dl1, s.z_t), fr_mul(c5, s2[ell + 2])));   // G_t
        fr_to_bytes(frow + 32 * (size_t)(n + 2), fr_sub(fr_mul(dl3, s.z_u), fr_mul(c5, s2[ell + 3])));   // G_u
    }
    // ---- variable coefficients: R | S | T | U | M | proof points | D ----
    for (uint32_t i = 0; i < ell; i++) {
        fr_to_bytes(vrow + 32 * (size_t)i, fr_neg(fr_mul(rho8, a[i])));
        fr_to_bytes(vrow + 32 * (size_t)(ell + i), fr_neg(fr_mul(rho9, a[i])));
        fr_to_bytes(vrow + 32 * (size_t)(2 * ell + i), fr_neg(fr_mul(c6, s2[i])));
        fr_to_bytes(vrow + 32 * (size_t)(3 * ell + i), fr_neg(fr_mul(c7, s2[i])));
    }
    fr_to_bytes(vrow + 32 * (size_t)(4 * ell), fr_neg(fr_mul(rho1, s.alpha_sp)));                         // M
    uint8_t* P = vrow + 32 * (size_t)NI;
    HFr a5 = fr_mul(rho5, alpha_msm);
    fr_to_bytes(P + 32 * L.A, fr_sub(a5, rho1));
    fr_to_bytes(P + 32 * L.T1, fr_sub(a5, fr_mul(dl1, alpha_ss)));
    fr_to_bytes(P + 32 * L.T2, fr_sub(fr_mul(rho6, alpha_msm), fr_mul(dl2, alpha_ss)));
    fr_to_bytes(P + 32 * L.U1, fr_sub(a5, fr_mul(dl3, alpha_ss)));
    fr_to_bytes(P + 32 * L.U2, fr_sub(fr_mul(rho7, alpha_msm), fr_mul(dl4, alpha_ss)));
    fr_to_bytes(P + 32 * L.R, fr_add(rho8, fr_mul(dl2, s.z_k)));
    fr_to_bytes(P + 32 * L.S, fr_add(rho9, fr_mul(dl4, s.z_k)));
    fr_to_bytes(P + 32 * L.B, rho1);
    fr_to_bytes(P + 32 * L.C, fr_mul(rho2, alpha_ipa));
    fr_to_bytes(P + 32 * L.Bc, rho2);
    fr_to_bytes(P + 32 * L.Bd, rho3);
    for (uint32_t j = 0; j < lg; j++) {
        fr_to_bytes(P + 32 * (L.LC + j), fr_mul(rho2, gam[j]));  fr_to_bytes(P + 32 * (L.RC + j), fr_mul(rho2, gam_inv[j]));
        fr_to_bytes(P + 32 * (L.LD + j), fr_mul(rho3, gam[j]));  fr_to_bytes(P + 32 * (L.RD + j), fr_mul(rho3, gam_inv[j]));
        fr_to_bytes(P + 32 * (L.LA + j), fr_mul(rho5, gam2[j])); fr_to_bytes(P + 32 * (L.RA + j), fr_mul(rho5, gam2_inv[j]));
        fr_to_bytes(P + 32 * (L.LT + j), fr_mul(rho6, gam2[j])); fr_to_bytes(P + 32 * (L.RT + j), fr_mul(rho6, gam2_inv[j]));
        fr_to_bytes(P + 32 * (L.LU + j), fr_mul(rho7, gam2[j])); fr_to_bytes(P + 32 * (L.RU + j), fr_mul(rho7, gam2_inv[j]));
    }
    fr_to_bytes(P + 32 * L.A1, fr_neg(dl1)); fr_to_bytes(P + 32 * L.A2, fr_neg(dl2));
    fr_to_bytes(P + 32 * L.B1, fr_neg(dl3)); fr_to_bytes(P + 32 * L.B2, fr_neg(dl4));
    fr_to_bytes(P + 32 * L.Ba, rho5); fr_to_bytes(P + 32 * L.Bt, rho6); fr_to_bytes(P + 32 * L.Bu, rho7);
    fr_to_bytes(vrow + 32 * (size_t)(NV - 1), fr_mul(rho3, alpha_ipa));                                  // D
}

// thread = proof (warp = 0), or WARP = proof (warp = 1: launched over 32 threads per proof; the 32 lanes run the same
// code on the same data in lock-step - identical loads, identical stores - and share the Keccak permutations)
struct VerifyPhase1 {
    static constexpr const char* kName = "VerifyPhase1";
    VShape sh; Layout L; VBuffers vb; uint64_t b0; uint32_t warp;
    CPG_HD void operator()(uint64_t t) const { verify_phase1(sh, L, vb, (size_t)(b0 + (warp ? t / 32 : t)), warp); }
};
struct VerifyPhase2 {
    static constexpr const char* kName = "VerifyPhase2";
    VShape sh; Layout L; VBuffers vb; uint64_t b0; uint32_t warp;
    CPG_HD void operator()(uint64_t t) const { verify_phase2(sh, L, vb, (size_t)(b0 + (warp ? t / 32 : t)), warp); }
};

struct MerlinScript {             // the whole script on one thread (cpg_merlin_script with on_device = 1) or on one warp in lock-step (= 2)
    static constexpr const char* kName = "MerlinScript";
    const uint8_t* script; size_t len; uint8_t* out; size_t cap; uint64_t* out_len; uint32_t warp;
    CPG_HD void operator()(uint64_t) const { *out_len = (uint64_t)cpgh::merlin_run_script(script, len, out, cap, warp); }
};

struct VerifyDerived {            // thread = proof: D = gh + B, A' = A + T_1 + U_1, their encodings
    static constexpr const char* kName = "VerifyDerived";
    Aff* bases;                   // [B][NV]; slot NV-1 receives D
    uint64_t NV;
    uint32_t ell, idxA, idxT1, idxU1, idxB;
    const Jac* gh;                // [B]: -beta^-1 G_sum + alpha H_sum (2-base fixed MSM)
    uint8_t* out;                 // [B][96]: compress(D) | compress(A')
    uint8_t* t0_inf;              // [B]: 1 if vec_T[0] is the identity
    CPG_HD void operator()(uint64_t b) const {
        Aff* row = bases + b * NV;
        Aff da = jac_to_aff(jac_add_mixed(gh[b], row[idxB]));
        row[NV - 1] = da;
        aff_compress(da, out + 96 * b);
        Jac ap = jac_add_mixed(jac_add_mixed(to_jac(row[idxA]), row[idxT1]), row[idxU1]);
        aff_compress(jac_to_aff(ap), out + 96 * b + 48);
        t0_inf[b] = is_inf(row[2 * (uint64_t)ell]) ? 1 : 0;
    }
};

struct AddFixedAndTest {          // thread = proof: verdict = !reject && (var + fixed == identity)
    static constexpr const char* kName = "AddFixedAndTest";
    const Jac* a; const Jac* b; const uint8_t* reject; uint8_t* ok;
    CPG_HD void operator()(uint64_t t) const { ok[t] = (!reject[t] && is_inf(jac_add(a[t], b[t]))) ? 1 : 0; }
};

// ---- cross-proof aggregation -------------------------------------------------------
// The 12 batching weights of a proof are drawn per proof (secret- and lane-keyed), so the SUM of the
// relations of G proofs is itself a random linear combination of all 12 G checks: a group of G proofs
// is accepted by ONE MSM over G*NV variable bases (wider windows: ~40 % fewer Fq products per proof at
// G = 8..32) plus ONE fixed-base MSM over the summed CRS coefficients.  A group that fails (or holds a
// structurally rejected proof) is re-checked proof by proof, so per-proof verdicts stay exact.
struct GroupSumFixed {            // thread = (group, i < NF): sum of the G proofs' coefficient of CRS base i
    static constexpr const char* kName = "GroupSumFixed";
    uint32_t G, NF; const uint8_t* fs; uint8_t* out;
    CPG_HD void operator()(uint64_t t) const {
        uint64_t g = t / NF; uint32_t i = (uint32_t)(t % NF);
        HFr acc = cpgh::fr_zero();
        for (uint32_t k = 0; k < G; k++) {
            HFr v;                                                       // canonical integers < r: add mod r as they are
            cpgh::copy32(v.l, fs + ((g * G + k) * (uint64_t)NF + i) * 32);
            acc = cpgh::fr_add(acc, v);
        }
        cpgh::copy32(out + t * 32, acc.l);
    }
};
struct GroupTest {                // thread = group: provisional verdict of its G proofs
    static constexpr const char* kName = "GroupTest";
    uint32_t G; const Jac* a; const Jac* b; const uint8_t* reject; uint8_t* gok; uint8_t* ok;
    CPG_HD void operator()(uint64_t g) const {
        bool fine = is_inf(jac_add(a[g], b[g]));
        for (uint32_t k = 0; k < G; k++) fine = fine && !reject[g * G + k];
        gok[g] = fine ? 1 : 0;
        for (uint32_t k = 0; k < G; k++) ok[g * G + k] = fine ? 1 : 0;
    }
};
struct Row16 { uint32_t w[4]; };
struct GatherRows {               // thread = (k, 16-byte word j): dst[k] = src[idx[k]]
    static constexpr const char* kName = "GatherRows";
    const uint32_t* idx; uint32_t words; const Row16* src; Row16* dst;
    CPG_HD void operator()(uint64_t t) const { uint64_t k = t / words, j = t % words; dst[t] = src[(uint64_t)idx[k] * words + j]; }
};
struct GatherMeta {               // thread = k: base offset (in points) and reject flag of proof idx[k]
    static constexpr const char* kName = "GatherMeta";
    const uint32_t* idx; uint32_t NV; const uint8_t* rej; uint32_t* off; uint8_t* rej_out;
    CPG_HD void operator()(uint64_t k) const { off[k] = idx[k] * NV; rej_out[k] = rej[idx[k]]; }
};
struct ScatterVerdicts {          // thread = k: ok[idx[k]] = per-proof verdict
    static constexpr const char* kName = "ScatterVerdicts";
    const uint32_t* idx; const Jac* a; const Jac* b; const uint8_t* rej; uint8_t* ok;
    CPG_HD void operator()(uint64_t k) const { ok[idx[k]] = (!rej[k] && is_inf(jac_add(a[k], b[k]))) ? 1 : 0; }
};

template <class F>
void parallel_for(int threads, size_t n, F f) {
    if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    int nt = (int)std::min<size_t>((size_t)threads, n);
    // items per grab: 8 for big batches; ONE when there are only a few items per thread (a 16-proof batch on 16 threads
    // was run by two of them, 8 transcripts each: 3.2 ms instead of 0.4)
    const size_t grab = std::max<size_t>(1, std::min<size_t>(8, n / ((size_t)nt * 4)));
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(grab); if (i >= n) break; for (size_t k = i; k < n && k < i + grab; k++) f(k); } });
    for (auto& th : pool) th.join();
}

// Temporarily route this thread's launches to another stream (see cur() in cpg_api.cu).
struct StreamScope {
#ifndef CPG_HOST_EMU
    cudaStream_t saved; bool saved_set, active;
    explicit StreamScope(cudaStream_t s) : saved(t_stream), saved_set(t_stream_set), active(s != nullptr) {
        if (active) { t_stream = s; t_stream_set = true; }
    }
    ~StreamScope() { if (active) { t_stream = saved; t_stream_set = saved_set; } }
#else
    explicit StreamScope(void*) {}
#endif
};

// A second stream for one of two INDEPENDENT latency chains.  fixed() and variable() are the two MSMs of a check (or of
// a prover round).  A handful of proofs (one shuffle per block) makes both of them latency chains - the 255 doublings of
// the variable-base Horner pass (1.3 ms) and the CRS table look-ups with their partial-sum tree (0.75 ms) - so with
// `small` the fixed-base one runs BESIDE the other on the side stream; big batches fill the GPU either way and keep
// the single stream.
struct SideLane {
#ifndef CPG_HOST_EMU
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    void destroy() {
        if (stream) cudaStreamDestroy(stream);
        for (cudaEvent_t e : {fork, join}) if (e) cudaEventDestroy(e);
        stream = nullptr; fork = join = nullptr;
    }
#else
    void destroy() {}
#endif
    template <class FX, class VA>
    int beside(bool small, FX fixed, VA variable) {
#ifndef CPG_HOST_EMU
        if (small) {
            if (!stream) {
                if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess)
                    return fail("side stream creation failed");
            }
            if (cudaEventRecord(fork, cur()) != cudaSuccess || cudaStreamWaitEvent(stream, fork, 0) != cudaSuccess)
                return fail("side stream fork failed");
            int rc = 0;
            {
                StreamScope scope(stream);
                rc = fixed();
            }
            if (!rc) rc = variable();
            if (cudaEventRecord(join, stream) != cudaSuccess || cudaStreamWaitEvent(cur(), join, 0) != cudaSuccess)
                if (!rc) rc = fail("side stream join failed");
            return rc;
        }
#else
        (void)small;
#endif
        if (int rc = variable()) return rc;
        return fixed();
    }
};

struct Verifier {
    // Sub-batch k owns three streams.  The transcript phases are one thread per proof: their duration is a LATENCY (a few
    // hundred dependent Keccak permutations), the same for 2048 proofs as for 8192, and they need next to no issue slots.
    // So they must never sit on the critical path of the integer-pipe kernels:
    //   tstreams[k] (highest priority)  VerifyPhase1(k): needs only the wire bytes, starts right after the upload
    //   streams[k]  (middle)            H2D(k) -> Decompress(k) -> [phase 1 done] -> D / A' -> VerifyPhase2(k)
    //   mstreams[k] (lowest)            the MSM check of sub-batch k
    // and the host enqueues EVERY sub-batch's first two rows before any check: all decompressions (and, under them, all
    // transcript work) are through before the last MSM starts, so no MSM ever waits for a latency-bound kernel.
#ifndef CPG_HOST_EMU
    cudaStream_t streams[8] = {}, tstreams[8] = {}, mstreams[8] = {};
    cudaEvent_t stream_done[8] = {}, ev_up[8] = {}, ev_p1[8] = {}, ev_p2[8] = {};

#else
    void *streams[8] = {}, *tstreams[8] = {}, *mstreams[8] = {};
#endif
    int nstreams = 8;
    SideLane side;                 // fixed-base MSM of a small sub-batch, beside its variable-base MSM
    // make the caller's stream wait for every sub-batch (its check stream ends the chain)
    int join_streams(int rc) {
#ifndef CPG_HOST_EMU
        for (int i = 0; i < 8; i++) {
            if (!mstreams[i]) continue;
            if (cudaEventRecord(stream_done[i], mstreams[i]) != cudaSuccess || cudaStreamWaitEvent(cur(), stream_done[i], 0) != cudaSuccess)
                if (!rc) rc = fail("cpg_verify_batch: stream join failed");
            if (cudaEventRecord(stream_done[i], streams[i]) != cudaSuccess || cudaStreamWaitEvent(cur(), stream_done[i], 0) != cudaSuccess)
                if (!rc) rc = fail("cpg_verify_batch: stream join failed");
        }
#endif
        return rc;
    }
    int fork_streams() {
#ifndef CPG_HOST_EMU
        // sub-batch streams start after whatever is already queued on the caller's stream
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);                 // numerically lower = more urgent
        const int mid = greatest < least ? greatest + 1 : least;
        for (int i = 0; i < 8 && i < (nstreams > 0 ? nstreams : 1); i++) {
            if (!streams[i]) {
                if (cudaStreamCreateWithPriority(&streams[i], cudaStreamNonBlocking, mid) != cudaSuccess ||
                    cudaStreamCreateWithPriority(&tstreams[i], cudaStreamNonBlocking, greatest) != cudaSuccess ||
                    cudaStreamCreateWithPriority(&mstreams[i], cudaStreamNonBlocking, least) != cudaSuccess)
                    return fail("cpg_verify_batch: stream creation failed");
                for (cudaEvent_t* e : {&stream_done[i], &ev_up[i], &ev_p1[i], &ev_p2[i]})
                    if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail("cpg_verify_batch: event creation failed");
            }
            if (cudaEventRecord(stream_done[i], cur()) != cudaSuccess || cudaStreamWaitEvent(streams[i], stream_done[i], 0) != cudaSuccess)
                return fail("cpg_verify_batch: stream fork failed");
        }
#endif
        return 0;
    }
    VShape sh;
    uint32_t nbl;
    size_t proof_len;             // 48 (M) + 1088 + 480 lg
    int threads;
    int transcript_mode = 2;      // 0 host threads, 1 one GPU thread per proof, 2 by batch size (cpg_verifier_set_transcript)
    int transcript_on_device = 1; // placement of the batch in flight (begin of cpg_verify_batch)
    // A CPU core runs one proof's Keccak chain ~20x faster than a lone GPU thread; the GPU runs thousands at once.
    // Mode 2 (default): every host thread gets at most one proof -> host; else the device: one WARP per proof while the
    // batch is a single sub-batch (nothing else to hide the one-thread-per-proof latency behind: phase 1 17.8 -> 9.5 ms at
    // B = 1024, the step 34.7 -> 30.3 ms), one THREAD per proof from two sub-batches on (the warp form only shares the
    // Keccak permutations - absorbing and the Fr loops still run on every lane - and costs 32x the issue slots: at
    // B = 4096 it is slower, 107 vs 70 ms).  Mode 3 forces warp per proof.
    bool device_transcript_for(size_t B) const { return transcript_mode == 1 || transcript_mode == 3 || (transcript_mode == 2 && B > (size_t)threads); }
    bool warp_transcript_for(size_t B) const { return transcript_mode == 3 || (transcript_mode == 2 && B <= 1024); }
    std::vector<uint8_t> crs48;   // vec_G | vec_H | H | G_t | G_u | G_sum | H_sum  (n + 5 points)
    Aff* d_crs = nullptr;         // n + 5 affine points
    uint8_t* d_crs48 = nullptr;   // the same as wire bytes (device transcript appends H)
    void* table = nullptr;        // fixed-base table over the first n + 3
    Shard shard;                  // world > 1: every proof's leaves split over ranks (cpg_verifier_create_sharded)
    std::vector<void*> shard_tables;   // table of each rank's block of the CRS bases (only this rank's unless the ranks are emulated)
    Jac* d_gather = nullptr; size_t gather_cap = 0;
    void* table_gh = nullptr;     // fixed-base table over G_sum, H_sum
    uint8_t secret[32];
    uint64_t calls = 0;           // batches checked so far (the weights' per-call nonce)
    int var_window = 0;
    uint32_t group = 1;           // proofs per aggregated check (1 = every proof on its own MSM)
    bool group_auto = false;      // re-pick `group` after every batch from the observed rate of failing proofs
    int group_window = 0;
    uint32_t cur_group = 1;       // group size of the batch in flight: `group`, halved until a batch of this size holds a whole group
    void begin_batch(size_t B) { cur_group = shard.on() ? 1 : group; while (cur_group > 1 && cur_group > B) cur_group >>= 1; }
    size_t rechecked = 0;         // proofs of the last batch that went through the per-proof fallback
    // device buffers of the current batch, kept (and grown on demand) between calls
    size_t cap = 0, lastB = 0;
    uint8_t *d_wire = nullptr, *d_psc = nullptr, *d_err = nullptr, *d_derived = nullptr, *d_t0 = nullptr, *d_vs = nullptr, *d_fs = nullptr,
            *d_ok = nullptr, *d_rej = nullptr, *d_chal = nullptr, *d_gok = nullptr, *d_gfs = nullptr;
    Aff* d_bases = nullptr; Jac *d_var = nullptr, *d_fix = nullptr, *d_gh = nullptr;
    VState* d_st = nullptr; HFr *d_a = nullptr, *d_tmp = nullptr;
    // pinned host staging (grown on demand)
    size_t hcap = 0;
    uint8_t *h_wire = nullptr, *h_psc = nullptr;
    // cache of decompressed tracker points (cpg_verifier_set_cache; msm.cuh "PointCache"); off while pc.lg == 0
    PointCache pc = {};
    uint8_t* d_role = nullptr; uint32_t *d_ref = nullptr, *d_todo = nullptr, *d_count = nullptr;
    unsigned long long launch_seq = 0;
    uint64_t cache_total[3] = {0, 0, 0};   // lookups, lookups served from the table, slots claimed - since creation / cpg_verifier_cache_reset
    void cache_free() {
        for (void* q : {(void*)pc.tag, (void*)pc.owner, (void*)pc.ready, (void*)pc.key, (void*)pc.val, (void*)pc.verr, (void*)pc.stats}) cpg_free(q);
        pc = PointCache{};
    }
    int cache_clear() {
        if (!pc.lg) return 0;
        const size_t slots = (size_t)1 << pc.lg;
        if (int rc = cpg_memset(pc.tag, 0, slots * 8)) return rc;
        if (int rc = cpg_memset(pc.ready, 0, slots * 4)) return rc;
        return cpg_memset(pc.stats, 0, 3 * 4);
    }
    // before a batch: fold the device counters into the totals; start over with an empty table once it is half full
    int cache_begin_batch() {
        if (!pc.lg) return 0;
        uint32_t st[3];
        if (int rc = cpg_d2h(st, pc.stats, sizeof st)) return rc;
        cache_total[0] += st[1]; cache_total[1] += st[2]; cache_total[2] += st[0];
        cache_slots_used += st[0];
        if (int rc = cpg_memset(pc.stats, 0, 3 * 4)) return rc;
        if (cache_slots_used > ((size_t)1 << pc.lg) / 2) { cache_slots_used = 0; return cache_clear(); }
        return 0;
    }
    size_t cache_slots_used = 0;

    std::vector<void*> all() { return {d_wire, d_psc, d_err, d_derived, d_t0, d_vs, d_fs, d_ok, d_rej, d_chal, d_gok, d_gfs, d_bases, d_var, d_fix, d_gh, d_st, d_a, d_tmp, d_role, d_ref, d_todo, d_count}; }
    void release() {
        for (void* q : all()) cpg_free(q);
        d_wire = d_psc = d_err = d_derived = d_t0 = d_vs = d_fs = d_ok = d_rej = d_chal = d_gok = d_gfs = nullptr;
        d_bases = nullptr; d_var = d_fix = d_gh = nullptr; d_st = nullptr; d_a = d_tmp = nullptr;
        d_role = nullptr; d_ref = d_todo = d_count = nullptr;
        cap = 0;
    }
    int reserve(size_t B) {
        if (B > hcap) {
            cpg_host_free(h_wire); cpg_host_free(h_psc);
            h_wire = (uint8_t*)cpg_host_alloc(B * sh.NV * 48);
            h_psc = (uint8_t*)cpg_host_alloc(B * 7 * 32);
            if (!h_wire || !h_psc) { hcap = 0; return fail("cpg_verify_batch: pinned host allocation failed"); }
            hcap = B;
        }
        if (B <= cap) return 0;
        release();
        const size_t NV = sh.NV, NF = sh.NF;
        d_wire = (uint8_t*)cpg_malloc(B * NV * 48);      d_bases = (Aff*)cpg_malloc(sizeof(Aff) * B * NV);
        d_psc = (uint8_t*)cpg_malloc(B * 7 * 32);        d_err = (uint8_t*)cpg_malloc(B * NV);
        d_chal = (uint8_t*)cpg_malloc(B * 64);           d_derived = (uint8_t*)cpg_malloc(B * 96);
        d_t0 = (uint8_t*)cpg_malloc(B);                  d_rej = (uint8_t*)cpg_malloc(B);
        d_vs = (uint8_t*)cpg_malloc(B * NV * 32);        d_fs = (uint8_t*)cpg_malloc(B * NF * 32);
        d_var = (Jac*)cpg_malloc(sizeof(Jac) * B);       d_fix = (Jac*)cpg_malloc(sizeof(Jac) * B);
        d_gh = (Jac*)cpg_malloc(sizeof(Jac) * B);        d_ok = (uint8_t*)cpg_malloc(B);
        d_st = (VState*)cpg_malloc(sizeof(VState) * B);  d_a = (HFr*)cpg_malloc(sizeof(HFr) * B * sh.ell);
        d_tmp = (HFr*)cpg_malloc(sizeof(HFr) * B * 5 * sh.n);
        d_gok = (uint8_t*)cpg_malloc(B);                 d_gfs = (uint8_t*)cpg_malloc((B / 2 + 1) * NF * 32);
        d_role = (uint8_t*)cpg_malloc(B * NV);           d_ref = (uint32_t*)cpg_malloc(B * NV * 4);
        d_todo = (uint32_t*)cpg_malloc(B * NV * 4);      d_count = (uint32_t*)cpg_malloc(8 * 4);
        for (void* q : all()) if (!q) { release(); return fail("cpg_verify_batch: device allocation failed"); }
        cap = B;
        return 0;
    }
    VBuffers device_buffers() {
        VBuffers vb;
        vb.wire = d_wire; vb.psc = d_psc; vb.crs48 = d_crs48; vb.st = d_st; vb.a = d_a; vb.tmp = d_tmp; vb.chal = d_chal;
        vb.derived = d_derived; vb.err = d_err; vb.t0 = d_t0; vb.vs = d_vs; vb.fs = d_fs; vb.reject = d_rej;
        memcpy(vb.secret, secret, 32);
        vb.nonce = ++calls;
        return vb;
    }
    // ---- the device side of proofs [b0, b0 + nb), in launch order (wire bytes already resident) ----
    // `k` = which of the 8 work-list counters this sub-batch uses (its stream index)
    int device_decode(size_t b0, size_t nb, size_t k = 0) {
        if (!pc.lg) return cpg_g1_decompress(d_wire + b0 * sh.NV * 48, nb * sh.NV, 0, d_bases + b0 * sh.NV, d_err + b0 * sh.NV);
        // cached path: the 4 ell tracker points of every proof go through the table, M and the proof's own points are
        // decoded as they are (unique per proof); one compact work list holds everything that needs a square root
        if (nb * (size_t)sh.NV >= 0xffffffffULL) return fail("cpg_verify_batch: sub-batch too large for the point cache");
        const unsigned long long id = ++launch_seq;
        const PointSel sel{sh.NV, 4 * sh.ell};
        const uint32_t* in = (const uint32_t*)(d_wire + b0 * sh.NV * 48);
        uint8_t* role = d_role + b0 * sh.NV; uint32_t* ref = d_ref + b0 * sh.NV; uint32_t* todo = d_todo + b0 * sh.NV; uint32_t* cnt = d_count + k;
        Aff* out = d_bases + b0 * sh.NV; uint8_t* err = d_err + b0 * sh.NV;
        if (int rc = cpg_memset(cnt, 0, 4)) return rc;
        if (int rc = launch(CacheClaim{pc, sel, id, in, role, ref}, nb * sel.cached)) return rc;
        if (int rc = launch(CacheResolve{pc, sel, id, 1u, sh.NV - 1, in, role, ref, todo, cnt}, nb * (size_t)(sh.NV - 1))) return rc;
        if (int rc = launch_decomp(DecompressList{pc, 1u, sel.cached, sh.NV, (const uint8_t*)in, todo, cnt, role, ref, out, err}, nb * (size_t)(sh.NV - 1))) return rc;
        return launch(CacheFill{pc, sel, role, ref, out, err}, nb * sel.cached);
    }
    int device_derive(size_t b0, size_t nb, const Layout& L) {
        if (int rc = cpg_g1_msm_fixed_batched(table_gh, d_chal + b0 * 64, nb, 0, d_gh + b0)) return rc;
        VerifyDerived vd{d_bases + b0 * sh.NV, sh.NV, sh.ell, sh.NI + L.A, sh.NI + L.T1, sh.NI + L.U1, sh.NI + L.B, d_gh + b0,
                         d_derived + b0 * 96, d_t0 + b0};
        return launch(vd, nb);
    }
    int device_check_each(size_t b0, size_t nb) {
        if (!nb) return 0;
        if (shard.on()) {
            // each rank sums its block of the NV variable bases and of the NF CRS bases of every proof; ONE all-gather
            // of 2 partial sums per proof, then the test on every rank (identical verdicts)
            const size_t cnt = 2 * nb;
            if ((size_t)shard.world * cnt > gather_cap) {
                cpg_free(d_gather);
                gather_cap = (size_t)shard.world * cnt;
                d_gather = (Jac*)cpg_malloc(sizeof(Jac) * gather_cap);
                if (!d_gather) { gather_cap = 0; return fail("cpg_verify_batch: device allocation failed"); }
            }
            for (int rk = shard.first(); rk < shard.last(); rk++) {
                Jac* slot = d_gather + (size_t)rk * cnt;
                size_t lo, hi;
                comm_block(sh.NV, rk, shard.world, &lo, &hi);
                if (int rc = msm_batched_impl(d_bases + b0 * sh.NV + lo, sh.NV, nullptr, d_vs + b0 * sh.NV * 32, nb, hi - lo, var_window, slot, 0, 0, sh.NV, lo)) return rc;
                comm_block(sh.NF, rk, shard.world, &lo, &hi);
                if (int rc = msm_fixed_impl(shard_tables[rk], d_fs + b0 * sh.NF * 32, nb, 0, slot + nb, sh.NF, lo)) return rc;
            }
            if (int rc = shard_exchange(shard, d_gather, cnt)) return rc;
            if (int rc = launch_occ(SumRanks{(uint32_t)shard.world, cnt, nb, d_gather, d_var + b0, d_fix + b0}, cnt)) return rc;
            return launch(AddFixedAndTest{d_var + b0, d_fix + b0, d_rej + b0, d_ok + b0}, nb);
        }
        if (int rc = side.beside(nb <= 64,
                            [&]() { return cpg_g1_msm_fixed_batched(table, d_fs + b0 * sh.NF * 32, nb, 0, d_fix + b0); },
                            [&]() { return cpg_g1_msm_batched(d_bases + b0 * sh.NV, sh.NV, d_vs + b0 * sh.NV * 32, nb, sh.NV, var_window, d_var + b0); })) return rc;
        return launch(AddFixedAndTest{d_var + b0, d_fix + b0, d_rej + b0, d_ok + b0}, nb);
    }
    // proofs [b0, b0 + nb), b0 a multiple of `group`: whole groups through one aggregated check each
    // (provisional verdicts; recheck_failed_groups settles the failing ones), the tail proof by proof
    int device_check(size_t b0, size_t nb) {
        const size_t G = cur_group;
        const size_t ng = G > 1 ? nb / G : 0, g0 = G > 1 ? b0 / G : 0;
        if (ng) {
            if (int rc = side.beside(ng <= 64,
                                [&]() {
                                    if (int r = launch(GroupSumFixed{(uint32_t)G, sh.NF, d_fs + b0 * sh.NF * 32, d_gfs + g0 * sh.NF * 32}, ng * sh.NF)) return r;
                                    return cpg_g1_msm_fixed_batched(table, d_gfs + g0 * sh.NF * 32, ng, 0, d_fix + g0);
                                },
                                [&]() { return cpg_g1_msm_batched(d_bases + b0 * sh.NV, G * sh.NV, d_vs + b0 * sh.NV * 32, ng, G * sh.NV, group_window, d_var + g0); })) return rc;
            if (int rc = launch(GroupTest{(uint32_t)G, d_var + g0, d_fix + g0, d_rej + b0, d_gok + g0, d_ok + b0}, ng)) return rc;
        }
        // the tail's scratch points share d_var/d_fix with the group results: indices >= b0 + ng*G > g0 + ng
        return device_check_each(b0 + ng * G, nb - ng * G);
    }
    // Group size for the next batch.  Cost per proof relative to per-proof MSMs: the aggregated MSM over
    // G*NV terms (measured at n = 128: 1, .85, .75, .68, .62, .60, .58 for G = 1..64) plus a per-proof
    // re-check with the probability that the proof's group holds a bad proof, 1 - (1 - p)^G.
    void adapt_group(double p_bad) {
        static const double agg[7] = {1.0, 0.85, 0.75, 0.68, 0.62, 0.60, 0.58};
        double best = 1e9; uint32_t bg = 1;
        for (int k = 0; k < 7; k++) {
            double G = (double)(1u << k);
            double cost = agg[k] + (k ? 1.0 - std::pow(1.0 - p_bad, G) : 0.0);
            if (cost < best - 1e-9) { best = cost; bg = 1u << k; }
        }
        group = bg < 2 ? 2 : bg;  // keep sampling the failure rate: at G = 1 nothing would be observed
    }
    // after every sub-batch has been joined: per-proof MSMs for the proofs of the failing groups
    int recheck_failed_groups(size_t B) {
        rechecked = 0;
        const size_t G = cur_group, ng = G > 1 ? B / G : 0;
        if (!ng) return 0;
        std::vector<uint8_t> gok(ng);
        if (int rc = cpg_d2h(gok.data(), d_gok, ng)) return rc;
        std::vector<uint32_t> idx;
        for (size_t g = 0; g < ng; g++) if (!gok[g]) for (size_t k = 0; k < G; k++) idx.push_back((uint32_t)(g * G + k));
        const size_t K = idx.size();
        if (group_auto) adapt_group((double)(K / G) / (double)(ng * G));
        if (!K) return 0;
        rechecked = K;
        Scratch sc;
        uint32_t* d_idx = sc.get<uint32_t>(K); uint32_t* d_off = sc.get<uint32_t>(K);
        uint8_t* vs2 = sc.get<uint8_t>(K * sh.NV * 32); uint8_t* fs2 = sc.get<uint8_t>(K * sh.NF * 32); uint8_t* rej2 = sc.get<uint8_t>(K);
        Jac* var2 = sc.get<Jac>(K); Jac* fix2 = sc.get<Jac>(K);
        if (!d_idx || !d_off || !vs2 || !fs2 || !rej2 || !var2 || !fix2) return fail("cpg_verify_batch: scratch allocation failed");
        if (int rc = cpg_h2d(d_idx, idx.data(), K * 4)) return rc;
        if (int rc = launch(GatherRows{d_idx, sh.NV * 2, (const Row16*)d_vs, (Row16*)vs2}, K * sh.NV * 2)) return rc;
        if (int rc = launch(GatherRows{d_idx, sh.NF * 2, (const Row16*)d_fs, (Row16*)fs2}, K * sh.NF * 2)) return rc;
        if (int rc = launch(GatherMeta{d_idx, sh.NV, d_rej, d_off, rej2}, K)) return rc;
        if (int rc = cpg_g1_msm_batched_off(d_bases, d_off, vs2, K, sh.NV, var_window, var2)) return rc;
        if (int rc = cpg_g1_msm_fixed_batched(table, fs2, K, 0, fix2)) return rc;
        if (int rc = launch(ScatterVerdicts{d_idx, var2, fix2, rej2, d_ok}, K)) return rc;
        return cpg_sync();                                               // idx (host) was read by an async copy
    }
    // Transcript on the device: nothing crosses PCIe between the stages.  The batch is cut into sub-batches on the
    // three-stream scheme above.  `upload` = also copy each sub-batch's wire bytes from the pinned staging buffers first;
    // `stage` (optional) fills those buffers for proofs [b0, b0 + nb) right before their copy is enqueued, so the host
    // stages sub-batch k + 1 while the GPU works on sub-batch k.
    int device_all(size_t B, bool upload, const std::function<void(size_t, size_t)>& stage = nullptr) {
        const Layout L(sh.lg);
        VBuffers vb = device_buffers();
        size_t S = nstreams > 0 ? (size_t)nstreams : 1;
        while (S > 1 && B < 1024 * S) S--;                   // sub-batches of >= 1024 proofs (measured: 8 x 1024 beats 4 x 2048 by 2.4 %)
        begin_batch(B);
        size_t align = cur_group > 32 ? cur_group : 32;     // warps of BucketAccumulate hold 32 MSMs; groups do not straddle sub-batches
        size_t per = ((B + S - 1) / S + align - 1) / align * align;
        int rc = 0;
        size_t si = 0;
#ifndef CPG_HOST_EMU
        const uint32_t wp = warp_transcript_for(B) ? 1u : 0u;
#else
        const uint32_t wp = 0;
#endif
        // rows 1 and 2 of every sub-batch
        for (size_t b0 = 0; b0 < B && !rc; b0 += per, si++) {
            const size_t nb = B - b0 < per ? B - b0 : per, k = si % 8;
            StreamScope scope(streams[k]);
            if (upload && stage) stage(b0, nb);
            if (upload) {
                rc = cpg_h2d(d_wire + b0 * sh.NV * 48, h_wire + b0 * sh.NV * 48, nb * sh.NV * 48);
                if (!rc) rc = cpg_h2d(d_psc + b0 * 224, h_psc + b0 * 224, nb * 224);
            }
#ifndef CPG_HOST_EMU
            if (!rc && (cudaEventRecord(ev_up[k], cur()) != cudaSuccess || cudaStreamWaitEvent(tstreams[k], ev_up[k], 0) != cudaSuccess)) rc = fail("cpg_verify_batch: stream fork failed");
#endif
            if (!rc) {
                StreamScope tscope(tstreams[k]);
                rc = launch<64>(VerifyPhase1{sh, L, vb, b0, wp}, wp ? nb * 32 : nb);
#ifndef CPG_HOST_EMU
                if (!rc && cudaEventRecord(ev_p1[k], cur()) != cudaSuccess) rc = fail("cpg_verify_batch: event record failed");
#endif
            }
            if (!rc) rc = device_decode(b0, nb, k);
#ifndef CPG_HOST_EMU
            if (!rc && cudaStreamWaitEvent(cur(), ev_p1[k], 0) != cudaSuccess) rc = fail("cpg_verify_batch: stream join failed");
#endif
            if (!rc) rc = device_derive(b0, nb, L);
            if (!rc) rc = launch<64>(VerifyPhase2{sh, L, vb, b0, wp}, wp ? nb * 32 : nb);
#ifndef CPG_HOST_EMU
            if (!rc && (cudaEventRecord(ev_p2[k], cur()) != cudaSuccess || cudaStreamWaitEvent(mstreams[k], ev_p2[k], 0) != cudaSuccess)) rc = fail("cpg_verify_batch: stream fork failed");
#endif
        }
        // row 3: the checks
        si = 0;
        for (size_t b0 = 0; b0 < B && !rc; b0 += per, si++) {
            const size_t nb = B - b0 < per ? B - b0 : per, k = si % 8;
            StreamScope scope(mstreams[k]);
            rc = device_check(b0, nb);
        }
        rc = join_streams(rc);
        if (!rc) rc = recheck_failed_groups(B);
        return rc;
    }
};

// 32 bytes for the per-process batching secret: getrandom(2), else /dev/urandom; CPG_TEST_NO_ENTROPY makes both
// fail so that the refusal path can be tested.  Never a constant (the reference draws fresh random weights per check,
// cp/msm_accumulator.py:43).
bool os_random_bytes(uint8_t* out, size_t n) {
    if (getenv("CPG_TEST_NO_ENTROPY")) return false;
    size_t got = 0;
#if defined(__linux__)
    while (got < n) {
        ssize_t r = getrandom(out + got, n - got, 0);
        if (r <= 0) break;
        got += (size_t)r;
    }
    if (got == n) return true;
#endif
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return false;
    got = fread(out, 1, n, f);
    fclose(f);
    return got == n;
}

// Split one wire proof (after M) into its points (48 B each, in order) and its 7 scalars.
void split_proof(const uint8_t* p, uint32_t lg, uint8_t* points48, uint8_t* scalars32) {
    auto pts = [&](uint32_t k) { memcpy(points48, p, 48 * (size_t)k); p += 48 * (size_t)k; points48 += 48 * (size_t)k; };
    auto sc = [&](uint32_t k) { memcpy(scalars32, p, 32 * (size_t)k); p += 32 * (size_t)k; scalars32 += 32 * (size_t)k; };
    pts(9); sc(1); pts(2 + 4 * lg); sc(2); pts(4); sc(3); pts(3 + 6 * lg); sc(1);
}

}  // namespace

extern "C" {

/* The library's STROBE-128 / Merlin transcript (host_transcript.h) driven by a byte script; replaces
 * merlin_transcripts.MerlinTranscript (merlin_transcript.py:6-24) and Strobe128 (strobe.py:16-107).  on_device = 1 runs
 * the same code in a one-thread kernel (the placement the batched prover / verifier use per proof). */
int cpg_merlin_script(const uint8_t* script, size_t len, int on_device, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!script || !out || !out_len) return fail("cpg_merlin_script: null argument");
    *out_len = 0;
    size_t got;
    if (!on_device) {
        got = cpgh::merlin_run_script(script, len, out, out_cap);
    } else {
        NEED_INIT();
        Scratch sc;
        uint8_t* d_sc = sc.get<uint8_t>(len + 8); uint8_t* d_out = sc.get<uint8_t>(out_cap + 8); uint64_t* d_len = sc.get<uint64_t>(1);
        if (!d_sc || !d_out || !d_len) return fail("cpg_merlin_script: scratch allocation failed");
        if (int rc = cpg_h2d(d_sc, script, len)) return rc;
        if (int rc = cpg_sync()) return rc;                                  // pageable source
        const uint32_t wp = on_device == 2 ? 1u : 0u;
        if (int rc = launch(MerlinScript{d_sc, len, d_out, out_cap, d_len, wp}, wp ? 32 : 1)) return rc;
        uint64_t g = 0;
        if (int rc = cpg_d2h(&g, d_len, 8)) return rc;
        got = (size_t)g;
        if (got != (size_t)-1 && got) if (int rc = cpg_d2h(out, d_out, got)) return rc;
    }
    if (got == (size_t)-1) return fail("cpg_merlin_script: malformed script or output buffer too small");
    *out_len = got;
    return 0;
}

/* Stateful form for host callers that interleave appends and challenges (the reference's usage,
 * cp/curdleproofs_transcript.py:6-25): host-only, one Transcript per handle. */
void* cpg_merlin_new(const uint8_t* label, size_t n) {
    cpgh::Transcript* t = new cpgh::Transcript;
    memset(t, 0, sizeof *t);
    t->init_rt(label, n);
    return t;
}
void* cpg_merlin_clone(const void* h) { return h ? new cpgh::Transcript(*(const cpgh::Transcript*)h) : nullptr; }
int cpg_merlin_free(void* h) { delete (cpgh::Transcript*)h; return 0; }
int cpg_merlin_append(void* h, const uint8_t* label, size_t nl, const uint8_t* msg, size_t n) {
    if (!h) return fail("cpg_merlin_append: null transcript");
    ((cpgh::Transcript*)h)->append_rt(label, nl, msg, n);
    return 0;
}
int cpg_merlin_challenge(void* h, const uint8_t* label, size_t nl, uint8_t* out, size_t n) {
    if (!h) return fail("cpg_merlin_challenge: null transcript");
    ((cpgh::Transcript*)h)->challenge_bytes_rt(label, nl, out, n);
    return 0;
}

static void* verifier_create_impl(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads, bool sharded);
void* cpg_verifier_create(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads) {
    return verifier_create_impl(crs_bytes, ell, n_blinders, fixed_window, host_threads, false);
}
/* Every proof over ALL ranks of the communicator (BASELINE config 5's verify side): every rank calls this and every later
 * cpg_verify_batch with identical arguments and gets identical verdicts; each proof's MSM terms are split over the
 * ranks and the 2 partial sums per proof cross by one all-gather.  The batching secret is rank 0's. */
void* cpg_verifier_create_sharded(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads) {
    return verifier_create_impl(crs_bytes, ell, n_blinders, fixed_window, host_threads, true);
}
static void* verifier_create_impl(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window, int host_threads, bool sharded) {
    if (need_init()) return nullptr;
    size_t n = ell + n_blinders;
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    if (((size_t)1 << lg) != n || n_blinders != 4 || ell < 2 || lg > MAX_LG) { fail("cpg_verifier_create: need ell + 4 = 2^k, k <= 16"); return nullptr; }
    Verifier* v = new Verifier;
    v->sh.ell = (uint32_t)ell; v->nbl = (uint32_t)n_blinders; v->sh.n = (uint32_t)n; v->sh.lg = lg;
    v->sh.NP = 18 + 10 * lg; v->sh.NI = 4 * (uint32_t)ell + 1; v->sh.NV = v->sh.NI + v->sh.NP + 1; v->sh.NF = (uint32_t)n + 3;
    v->proof_len = 48 + 1088 + 480 * (size_t)lg;
    v->threads = host_threads > 0 ? host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    v->crs48.assign(crs_bytes, crs_bytes + 48 * (n + 5));
    // The batching weights are sound only while the prover cannot predict them: no entropy, no verifier.
    if (!os_random_bytes(v->secret, 32)) { delete v; fail("cpg_verifier_create: the OS gave no random bytes (getrandom and /dev/urandom both failed)"); return nullptr; }
    v->d_crs48 = (uint8_t*)cpg_malloc(48 * (n + 5));
    uint8_t* derr = (uint8_t*)cpg_malloc(n + 5);
    v->d_crs = (Aff*)cpg_malloc(sizeof(Aff) * (n + 5));
    std::vector<uint8_t> err(n + 5, 1);
    int rc = (!v->d_crs48 || !derr || !v->d_crs) ? fail("cpg_verifier_create: allocation failed") : 0;
    if (!rc) rc = cpg_h2d(v->d_crs48, crs_bytes, 48 * (n + 5));
    if (!rc) rc = cpg_g1_decompress(v->d_crs48, n + 5, 0, v->d_crs, derr);
    if (!rc) rc = cpg_d2h(err.data(), derr, n + 5);
    cpg_free(derr);
    if (!rc) for (uint8_t e : err) if (e) { rc = fail("cpg_verifier_create: CRS holds an invalid point encoding"); break; }
    if (sharded) v->shard = shard_now();
    if (!rc && v->shard.on()) {
        // all ranks must weigh the checks alike: rank 0's secret reaches the others through the communicator
        if (!v->shard.virt) {
            Scratch sc;
            uint8_t* d_all = sc.get<uint8_t>(32 * (size_t)v->shard.world);
            if (!d_all) rc = fail("cpg_verifier_create_sharded: scratch allocation failed");
            if (!rc) rc = cpg_h2d(d_all + 32 * (size_t)v->shard.rank, v->secret, 32);
            if (!rc) rc = cpg_sync();
            if (!rc) rc = comm_allgather(d_all + 32 * (size_t)v->shard.rank, d_all, 32);
            if (!rc) rc = cpg_d2h(v->secret, d_all, 32);
        }
        v->shard_tables.assign(v->shard.world, nullptr);
        for (int rk = v->shard.first(); rk < v->shard.last() && !rc; rk++) {
            size_t lo, hi;
            comm_block(n + 3, rk, v->shard.world, &lo, &hi);
            if (hi == lo) { rc = fail("cpg_verifier_create_sharded: more ranks than CRS bases"); break; }
            v->shard_tables[rk] = cpg_fixed_table_create(v->d_crs + lo, hi - lo, fixed_window > 0 ? fixed_window : 12);
            if (!v->shard_tables[rk]) rc = 1;
        }
    } else if (!rc) { v->table = cpg_fixed_table_create(v->d_crs, n + 3, fixed_window > 0 ? fixed_window : 12); if (!v->table) rc = 1; }
    if (!rc) { v->table_gh = cpg_fixed_table_create(v->d_crs + (n + 3), 2, 8); if (!v->table_gh) rc = 1; }
    if (rc) { cpg_fixed_table_free(v->table); cpg_free(v->d_crs); cpg_free(v->d_crs48); delete v; return nullptr; }
    return v;
}

int cpg_verifier_free(void* handle) {
    if (!handle) return 0;
    Verifier* v = (Verifier*)handle;
    v->release();
#ifndef CPG_HOST_EMU
    for (int i = 0; i < 8; i++) {
        for (cudaStream_t st : {v->streams[i], v->tstreams[i], v->mstreams[i]}) if (st) cudaStreamDestroy(st);
        for (cudaEvent_t e : {v->stream_done[i], v->ev_up[i], v->ev_p1[i], v->ev_p2[i]}) if (e) cudaEventDestroy(e);
    }
#endif
    v->side.destroy();
    cpg_host_free(v->h_wire); cpg_host_free(v->h_psc);
    v->cache_free();
    cpg_fixed_table_free(v->table);
    cpg_fixed_table_free(v->table_gh);
    for (void* t : v->shard_tables) cpg_fixed_table_free(t);
    cpg_free(v->d_gather);
    cpg_free(v->d_crs);
    cpg_free(v->d_crs48);
    delete v;
    return 0;
}

size_t cpg_verifier_proof_bytes(const void* handle) { return handle ? ((const Verifier*)handle)->proof_len : 0; }
size_t cpg_verifier_input_bytes(const void* handle) { return handle ? (size_t)(((const Verifier*)handle)->sh.NI - 1) * 48 : 0; }
int cpg_verifier_set_window(void* handle, int var_window) { if (!handle) return 1; ((Verifier*)handle)->var_window = var_window; return 0; }
int cpg_verifier_set_streams(void* handle, int nstreams) {
    if (!handle || nstreams < 1 || nstreams > 8) return fail("cpg_verifier_set_streams: 1..8");
    ((Verifier*)handle)->nstreams = nstreams;
    return 0;
}
int cpg_verifier_set_group(void* handle, int group, int group_window) {
    if (!handle) return fail("cpg_verifier_set_group: null verifier");
    if (group < 0 || group > 4096 || (group & (group - 1))) return fail("cpg_verifier_set_group: the group size must be 0 (adaptive) or a power of two in [1, 4096]");
    if (group_window < 0 || group_window > 16) return fail("cpg_verifier_set_group: bad window");
    Verifier& v = *(Verifier*)handle;
    v.group_auto = group == 0;
    v.group = group ? (uint32_t)group : 16; v.group_window = group_window;
    return 0;
}
/* Device-resident cache of decompressed tracker points (2^log2_slots entries of 160 B; 0 = off, the default). */
int cpg_verifier_set_cache(void* handle, int log2_slots) {
    if (!handle) return fail("cpg_verifier_set_cache: null verifier");
    if (log2_slots != 0 && (log2_slots < 10 || log2_slots > 28)) return fail("cpg_verifier_set_cache: 0 (off) or 10..28");
    Verifier& v = *(Verifier*)handle;
    if (int rc = cpg_sync()) return rc;
    v.cache_free();
    v.cache_slots_used = 0;
    if (!log2_slots) return 0;
    const size_t slots = (size_t)1 << log2_slots;
    PointCache pc{};
    pc.lg = (uint32_t)log2_slots; pc.max_probe = 64;
    pc.tag = (unsigned long long*)cpg_malloc(slots * 8); pc.owner = (unsigned long long*)cpg_malloc(slots * 8);
    pc.ready = (uint32_t*)cpg_malloc(slots * 4); pc.key = (uint32_t*)cpg_malloc(slots * 48);
    pc.val = (Aff*)cpg_malloc(slots * sizeof(Aff)); pc.verr = (uint8_t*)cpg_malloc(slots); pc.stats = (uint32_t*)cpg_malloc(3 * 4);
    v.pc = pc;
    if (!pc.tag || !pc.owner || !pc.ready || !pc.key || !pc.val || !pc.verr || !pc.stats) { v.cache_free(); return fail("cpg_verifier_set_cache: device allocation failed"); }
    if (int rc = cpg_memset(pc.owner, 0, slots * 8)) return rc;
    return v.cache_clear();
}
int cpg_verifier_cache_reset(void* handle) {
    if (!handle) return fail("cpg_verifier_cache_reset: null verifier");
    Verifier& v = *(Verifier*)handle;
    if (int rc = cpg_sync()) return rc;
    v.cache_slots_used = 0;
    v.cache_total[0] = v.cache_total[1] = v.cache_total[2] = 0;
    return v.cache_clear();
}
int cpg_verifier_cache_stats(void* handle, uint64_t* out3) {
    if (!handle || !out3) return fail("cpg_verifier_cache_stats: null argument");
    Verifier& v = *(Verifier*)handle;
    out3[0] = v.cache_total[0]; out3[1] = v.cache_total[1]; out3[2] = v.cache_total[2];
    if (!v.pc.lg) return 0;
    uint32_t st[3];
    if (int rc = cpg_d2h(st, v.pc.stats, sizeof st)) return rc;      // the batch in flight, not folded in yet
    out3[0] += st[1]; out3[1] += st[2]; out3[2] += st[0];
    return 0;
}
int cpg_verifier_group(const void* handle) { return handle ? (int)((const Verifier*)handle)->group : 0; }
size_t cpg_verifier_rechecked(const void* handle) { return handle ? ((const Verifier*)handle)->rechecked : 0; }
int cpg_verifier_set_transcript(void* handle, int mode) {
    if (!handle) return fail("cpg_verifier_set_transcript: null verifier");
    if (mode < 0 || mode > 3) return fail("cpg_verifier_set_transcript: 0 (host threads), 1 (GPU thread per proof), 2 (by batch size) or 3 (GPU warp per proof)");
    ((Verifier*)handle)->transcript_mode = mode;
    return 0;
}

/* inputs : [B][4*ell*48]   vec_R | vec_S | vec_T | vec_U   (tracker halves, whisk_interface.py:96-100)
 * proofs : [B][proof_len]  M | proof                        (WhiskShuffleProof.to_bytes, :57-61)
 * verdicts[b] = 1 iff the reference's IsValidWhiskShuffleProof would return True. */
int cpg_verify_batch(void* handle, const uint8_t* inputs, const uint8_t* proofs, size_t B, uint8_t* verdicts) {
    NEED_INIT();
    if (!handle) return fail("cpg_verify_batch: null verifier");
    if (!B) return 0;
    Verifier& v = *(Verifier*)handle;
    const VShape sh = v.sh;
    const uint32_t lg = sh.lg, NI = sh.NI, NV = sh.NV, NF = sh.NF;
    const Layout L(lg);
    const size_t in_len = (size_t)(NI - 1) * 48;
    if (int rc = v.reserve(B)) return rc;
    if (int rc = v.cache_begin_batch()) return rc;
    v.lastB = B;
    v.transcript_on_device = v.device_transcript_for(B) ? 1 : 0;

    // ---- stage wire points contiguously per proof: R|S|T|U|M|proof points|(slot for D) ----
    uint8_t* wire = v.h_wire;
    uint8_t* psc = v.h_psc;
    auto stage = [&](size_t b0, size_t nb) {
        parallel_for(v.threads, nb, [&](size_t i) {
            const size_t b = b0 + i;
            uint8_t* row = wire + b * (size_t)NV * 48;
            memcpy(row, inputs + b * in_len, in_len);
            const uint8_t* pr = proofs + b * v.proof_len;
            memcpy(row + in_len, pr, 48);                                   // M
            split_proof(pr + 48, lg, row + (size_t)NI * 48, psc + b * 7 * 32);
            memset(row + (size_t)(NV - 1) * 48, 0, 48);
            row[(size_t)(NV - 1) * 48] = 0xc0;                              // D slot: decodes to identity
        });
    };
    if (v.transcript_on_device) {
        if (int rc = v.fork_streams()) return rc;
        if (int rc = v.device_all(B, true, stage)) return rc;               // staging of sub-batch k+1 overlaps the GPU work on k
        return cpg_d2h(verdicts, v.d_ok, B);
    }
    stage(0, B);

    // ---- transcript on host threads ----
    if (int rc = cpg_h2d(v.d_wire, wire, (size_t)B * NV * 48)) return rc;
    if (int rc = cpg_h2d(v.d_psc, psc, (size_t)B * 7 * 32)) return rc;
    if (int rc = v.device_decode(0, B)) return rc;
    std::vector<VState> st(B);
    std::vector<HFr> a((size_t)B * sh.ell), tmp((size_t)B * 5 * sh.n);
    std::vector<uint8_t> chal((size_t)B * 64), derived((size_t)B * 96), err((size_t)B * NV), t0(B);
    std::vector<uint8_t> vs((size_t)B * NV * 32), fs((size_t)B * NF * 32), reject(B, 0);
    VBuffers vb;
    vb.wire = wire; vb.psc = psc; vb.crs48 = v.crs48.data(); vb.st = st.data(); vb.a = a.data(); vb.tmp = tmp.data(); vb.chal = chal.data();
    vb.derived = derived.data(); vb.err = err.data(); vb.t0 = t0.data(); vb.vs = vs.data(); vb.fs = fs.data(); vb.reject = reject.data();
    memcpy(vb.secret, v.secret, 32);
    vb.nonce = ++v.calls;
    parallel_for(v.threads, B, [&](size_t b) { verify_phase1(sh, L, vb, b); });     // overlaps the decompression kernel
    if (int rc = cpg_h2d(v.d_chal, chal.data(), chal.size())) return rc;
    if (int rc = v.device_derive(0, B, L)) return rc;
    if (int rc = cpg_d2h(derived.data(), v.d_derived, derived.size())) return rc;
    if (int rc = cpg_d2h(err.data(), v.d_err, err.size())) return rc;
    if (int rc = cpg_d2h(t0.data(), v.d_t0, B)) return rc;
    parallel_for(v.threads, B, [&](size_t b) { verify_phase2(sh, L, vb, b); });
    if (int rc = cpg_h2d(v.d_vs, vs.data(), vs.size())) return rc;
    if (int rc = cpg_h2d(v.d_fs, fs.data(), fs.size())) return rc;
    if (int rc = cpg_h2d(v.d_rej, reject.data(), B)) return rc;
    v.begin_batch(B);
    if (int rc = v.device_check(0, B)) return rc;
    if (int rc = v.recheck_failed_groups(B)) return rc;
    return cpg_d2h(verdicts, v.d_ok, B);
}

/* Re-runs the device side of the last cpg_verify_batch on its (still resident) wire bytes:
 * decompress -> [transcript kernels] -> D / A' -> per-proof MSM -> verdict.  With the transcript on
 * the host, the coefficients of the last call are reused.  Used to time the GPU path with inputs in HBM. */
int cpg_verify_replay_device(void* handle, uint8_t* verdicts) {
    NEED_INIT();
    if (!handle) return fail("cpg_verify_replay_device: null verifier");
    Verifier& v = *(Verifier*)handle;
    if (!v.lastB) return fail("cpg_verify_replay_device: no batch resident");
    if (int rc = v.cache_begin_batch()) return rc;
    if (v.transcript_on_device) {
        if (int rc = v.fork_streams()) return rc;
        if (int rc = v.device_all(v.lastB, false)) return rc;
    } else {
        const Layout L(v.sh.lg);
        if (int rc = v.device_decode(0, v.lastB)) return rc;
        if (int rc = v.device_derive(0, v.lastB, L)) return rc;
        v.begin_batch(v.lastB);
        if (int rc = v.device_check(0, v.lastB)) return rc;
        if (int rc = v.recheck_failed_groups(v.lastB)) return rc;
    }
    if (verdicts) return cpg_d2h(verdicts, v.d_ok, v.lastB);
    return 0;
}

}  // extern "C"

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

struct PShape { uint32_t ell, n, lg, NF, NR, rounds; };

// per-proof scalar state; the Fr vectors live in PBuffers::vec
struct PState {
    cpgh::Transcript tr;
    HFr k, alpha_sp, beta_sp, gprod, alpha_gp, beta_gp, beta_gp_inv, r_p, z;
    HFr alpha_ipa, beta_ipa, r_t, r_u, r_a, r_b, r_k, z_k, z_t, z_u, alpha_msm;
    uint32_t len;                 // current folded length of the running loop
};

// vectors per proof, each n entries: see PV_* below
enum { PV_A = 0, PV_APERM, PV_FACT, PV_C, PV_D, PV_X, PV_WG, PV_WGP, PV_W2, PV_ACOEF, PV_MCOEF, PV_BCOEF, PV_U, PV_COUNT };

struct PBuffers {
    const uint8_t* in48;          // [B][2 ell][48]  vec_R | vec_S wire bytes
    const uint8_t* tu48;          // [B][2 ell][48]  vec_T | vec_U wire bytes (round 0 output)
    const uint32_t* perm;         // [B][ell]
    const uint8_t* kbytes;        // [B][32]
    const uint8_t* rand;          // [B][NR][32]     m_bl(4) a_bl(2) c_bl(4) ipa_r(n) ipa_z(n-2) r_t r_u r_a r_b r_k msm_r(n)
    const uint8_t* crs48;         // CRS wire bytes
    PState* st;                   // [B]
    HFr* vec;                     // [PV_COUNT][n][B]  (batch-innermost, see FrVec)
    uint8_t* outs48;              // [B][NOUT][48]   compressed outputs, by output id
    uint8_t* fs;                  // [P_MAX_OUT][B][NF][32]  fixed-base coefficient rows of the round being prepared (output-major)
    uint8_t* vs;                  // [P_MAX_VAR][B][ell][32] variable-base coefficient rows
    uint8_t* proof;               // [B][proof_len]  wire proof (without M), filled at the end
    uint64_t B;
};

// A per-proof Fr vector.  Storage is batch-innermost, vec[which][i][b]: the 32 lanes of a warp (32
// consecutive proofs) touch 32 consecutive elements, instead of 32 rows that lie 13*n*32 bytes apart.
struct FrVec {
    HFr* p; uint64_t stride;
    CPG_HD HFr& operator[](uint32_t i) const { return p[(uint64_t)i * stride]; }
    CPG_HD FrVec operator+(uint32_t k) const { return FrVec{p + (uint64_t)k * stride, stride}; }
};
CPG_HD FrVec pvec(const PShape& sh, const PBuffers& pb, size_t b, int which) { return FrVec{pb.vec + (size_t)which * sh.n * pb.B + b, pb.B}; }
CPG_HD uint8_t* frow(const PShape& sh, const PBuffers& pb, size_t b, uint32_t o) { return pb.fs + ((size_t)o * pb.B + b) * (size_t)sh.NF * 32; }
CPG_HD uint8_t* vrow(const PShape& sh, const PBuffers& pb, size_t b, uint32_t o) { return pb.vs + ((size_t)o * pb.B + b) * (size_t)sh.ell * 32; }
CPG_HD HFr prand(const PBuffers& pb, const PShape& sh, size_t b, uint32_t i) {
    HFr v; cpgh::fr_from_bytes(&v, pb.rand + (b * sh.NR + i) * 32); return v;
}
CPG_HD void zero_rows(const PShape& sh, const PBuffers& pb, size_t b, uint32_t nfix, uint32_t nvar) {
    for (uint32_t o = 0; o < nfix; o++) cpgh::zero_bytes(frow(sh, pb, b, o), (size_t)sh.NF * 32);
    for (uint32_t o = 0; o < nvar; o++) cpgh::zero_bytes(vrow(sh, pb, b, o), (size_t)sh.ell * 32);
}
CPG_HD HFr ip(const FrVec& a, const FrVec& b, uint32_t n) {
    HFr acc = cpgh::fr_zero();
    for (uint32_t i = 0; i < n; i++) acc = cpgh::fr_add(acc, cpgh::fr_mul(a[i], b[i]));
    return acc;
}
// fixed-table index of leaf L of G_wb = vec_G | vec_H[:2] | G_t | G_u
CPG_HD uint32_t gwb_index(const PShape& sh, uint32_t L) { return L < sh.ell + 2 ? L : L + 3; }   // ell+2 -> n+1 (G_t), ell+3 -> n+2 (G_u)

// offsets into the rand array
struct PRand { uint32_t m_bl = 0, a_bl = 4, c_bl = 6, ipa_r = 10, ipa_z, r_t, r_u, r_a, r_b, r_k, msm_r, NR;
    CPG_HD explicit PRand(uint32_t n) { ipa_z = ipa_r + n; r_t = ipa_z + (n - 2); r_u = r_t + 1; r_a = r_u + 1; r_b = r_a + 1; r_k = r_b + 1; msm_r = r_k + 1; NR = msm_r + n; } };

// One round of one proof.  Rounds: 0 M | 1 A | 2 B | 3 C | 4 D,B_c,B_d | 5..4+lg IPA | 5+lg SameScalar |
// 6+lg A',B_a,B_t,B_u | 7+lg..6+2lg SameMSM | 7+2lg finish (assemble the wire proof).
CPG_HD void prove_step(const PShape& sh, const POut& O, const PBuffers& pb, uint32_t round, size_t b) {
    using namespace cpgh;
    const uint32_t ell = sh.ell, n = sh.n, lg = sh.lg;
    const PRand RO(n);
    PState& s = pb.st[b];
    const uint8_t* outs = pb.outs48 + b * (size_t)O.NOUT * 48;
    FrVec a = pvec(sh, pb, b, PV_A);
    FrVec aperm = pvec(sh, pb, b, PV_APERM);
    FrVec fact = pvec(sh, pb, b, PV_FACT);
    FrVec c = pvec(sh, pb, b, PV_C);
    FrVec d = pvec(sh, pb, b, PV_D);
    FrVec x = pvec(sh, pb, b, PV_X);
    FrVec wG = pvec(sh, pb, b, PV_WG);
    FrVec wGp = pvec(sh, pb, b, PV_WGP);
    FrVec w2 = pvec(sh, pb, b, PV_W2);
    FrVec Acoef = pvec(sh, pb, b, PV_ACOEF);
    FrVec Mcoef = pvec(sh, pb, b, PV_MCOEF);
    FrVec Bcoef = pvec(sh, pb, b, PV_BCOEF);
    FrVec uvec = pvec(sh, pb, b, PV_U);
    const uint32_t* perm = pb.perm + b * (size_t)ell;
    Transcript tr;
    if (round > 0) tr = s.tr;

    if (round == 0) {                                   // M = MSM(vec_G, sigma) + MSM(vec_H, m_bl)   (curdleproofs.py:310-319)
        fr_from_bytes(&s.k, pb.kbytes + b * 32);
        zero_rows(sh, pb, b, 1, 0);
        for (uint32_t i = 0; i < n; i++) Mcoef[i] = i < ell ? fr_from_u64(perm[i]) : prand(pb, sh, b, RO.m_bl + (i - ell));
        uint8_t* f = frow(sh, pb, b, 0);
        for (uint32_t i = 0; i < n; i++) fr_to_bytes(f + 32 * (size_t)i, Mcoef[i]);
        return;                                         // no transcript yet
    }
    if (round == 1) {                                   // step1 -> vec_a -> A                         (curdleproofs.py:65-77)
        tr.init("curdleproofs");
        const uint8_t* in = pb.in48 + b * (size_t)(2 * ell) * 48;
        const uint8_t* tu = pb.tu48 + b * (size_t)(2 * ell) * 48;
        for (uint32_t i = 0; i < 2 * ell; i++) tr.append_point("curdleproofs_step1", in + 48 * (size_t)i);
        for (uint32_t i = 0; i < 2 * ell; i++) tr.append_point("curdleproofs_step1", tu + 48 * (size_t)i);
        tr.append_point("curdleproofs_step1", outs + 48 * O.M);
        for (uint32_t i = 0; i < ell; i++) a[i] = tr.challenge("curdleproofs_vec_a");
        for (uint32_t i = 0; i < ell; i++) aperm[i] = a[perm[i]];
        aperm[ell] = prand(pb, sh, b, RO.a_bl); aperm[ell + 1] = prand(pb, sh, b, RO.a_bl + 1);   // a_bl; r_a' = a_bl | 0 0
        aperm[ell + 2] = fr_zero(); aperm[ell + 3] = fr_zero();
        zero_rows(sh, pb, b, 1, 0);
        uint8_t* f = frow(sh, pb, b, 0);
        for (uint32_t i = 0; i < n; i++) { Acoef[i] = aperm[i]; fr_to_bytes(f + 32 * (size_t)i, Acoef[i]); }
        s.tr = tr;
        return;
    }
    if (round == 2) {                                   // same_perm: alpha, beta, B                   (same_perm.py:43-55)
        tr.append_point("same_perm_step1", outs + 48 * O.A);
        tr.append_point("same_perm_step1", outs + 48 * O.M);
        for (uint32_t i = 0; i < ell; i++) tr.append_fr("same_perm_step1", a[i]);
        s.alpha_sp = tr.challenge("same_perm_alpha");
        s.beta_sp = tr.challenge("same_perm_beta");
        HFr g = fr_one();
        for (uint32_t i = 0; i < ell; i++) {
            fact[i] = fr_add(fr_add(aperm[i], fr_mul(fr_from_u64(perm[i]), s.alpha_sp)), s.beta_sp);
            g = fr_mul(g, fact[i]);
        }
        s.gprod = g;
        zero_rows(sh, pb, b, 1, 0);
        uint8_t* f = frow(sh, pb, b, 0);
        for (uint32_t i = 0; i < n; i++) {
            HFr t = fr_add(Acoef[i], fr_mul(s.alpha_sp, Mcoef[i]));
            if (i < ell) t = fr_add(t, s.beta_sp);
            Bcoef[i] = t;
            fr_to_bytes(f + 32 * (size_t)i, t);
        }
        s.tr = tr;
        return;
    }
    if (round == 3) {                                   // gprod step 1: alpha, C                      (grand_prod.py:44-54)
        tr.append_point("gprod_step1", outs + 48 * O.B);
        tr.append_fr("gprod_step1", s.gprod);
        s.alpha_gp = tr.challenge("gprod_alpha");
        c[0] = fr_one();
        for (uint32_t i = 0; i + 1 < ell; i++) c[i + 1] = fr_mul(c[i], fact[i]);
        for (uint32_t i = 0; i < 4; i++) c[ell + i] = prand(pb, sh, b, RO.c_bl + i);
        // b_bl = r_a' + alpha_sp m_bl ; rb_alpha = b_bl + alpha_gp ; r_p = <rb_alpha, c_bl>
        HFr rp = fr_zero();
        for (uint32_t i = 0; i < 4; i++) {
            HFr bbl = fr_add(aperm[ell + i], fr_mul(s.alpha_sp, Mcoef[ell + i]));
            d[ell + i] = fr_add(bbl, s.alpha_gp);        // park rb_alpha in d's blinder slots
            rp = fr_add(rp, fr_mul(d[ell + i], c[ell + i]));
        }
        s.r_p = rp;
        zero_rows(sh, pb, b, 1, 0);
        uint8_t* f = frow(sh, pb, b, 0);
        for (uint32_t i = 0; i < n; i++) fr_to_bytes(f + 32 * (size_t)i, c[i]);
        s.tr = tr;
        return;
    }
    if (round == 4) {                                   // gprod step 2: beta, D; IPA blinders, B_c, B_d (grand_prod.py:59-90, ipa.py:27-48,97-98)
        tr.append_point("gprod_step2", outs + 48 * O.C);
        tr.append_fr("gprod_step2", s.r_p);
        s.beta_gp = tr.challenge("gprod_beta");
        s.beta_gp_inv = fr_inv(s.beta_gp);
        // d_i = b_i beta^(i+1) - beta^i ; d_bl = beta^(ell+1) rb_alpha ; u_i = beta^-(i+1) (blinders beta^-(ell+1))
        HFr pw = fr_one(), ui = s.beta_gp_inv;
        for (uint32_t i = 0; i < ell; i++) {
            HFr pw1 = fr_mul(pw, s.beta_gp);
            d[i] = fr_sub(fr_mul(fact[i], pw1), pw);
            uvec[i] = ui;
            ui = fr_mul(ui, s.beta_gp_inv);
            pw = pw1;
        }
        HFr beta_l = pw, beta_l1 = fr_mul(pw, s.beta_gp);     // beta^ell, beta^(ell+1)
        for (uint32_t i = 0; i < 4; i++) { d[ell + i] = fr_mul(beta_l1, d[ell + i]); uvec[ell + i] = ui; }   // ui = beta^-(ell+1)
        s.z = fr_sub(fr_add(fr_mul(s.r_p, beta_l1), fr_mul(s.gprod, beta_l)), fr_one());
        // IPA blinders (ipa.py:27-48): r in x (scratch), z in w2 (scratch)
        FrVec r = x, zz = w2;
        for (uint32_t i = 0; i < n; i++) r[i] = prand(pb, sh, b, RO.ipa_r + i);
        for (uint32_t i = 0; i + 2 < n; i++) zz[i] = prand(pb, sh, b, RO.ipa_z + i);
        HFr omega = fr_add(ip(r, d, n), ip(zz, c, n - 2));
        HFr delta = ip(r, zz, n - 2);
        HFr inv_c = fr_inv(c[n - 2]);
        HFr t1 = fr_mul(r[n - 2], inv_c);
        HFr last_z = fr_mul(fr_sub(fr_mul(t1, omega), delta), fr_inv(fr_add(fr_neg(fr_mul(t1, c[n - 1])), r[n - 1])));
        HFr pen_z = fr_neg(fr_mul(inv_c, fr_add(fr_mul(last_z, c[n - 1]), omega)));
        zz[n - 2] = pen_z; zz[n - 1] = last_z;
        zero_rows(sh, pb, b, 3, 0);
        uint8_t* fD = frow(sh, pb, b, 0); uint8_t* fBc = frow(sh, pb, b, 1); uint8_t* fBd = frow(sh, pb, b, 2);
        for (uint32_t i = 0; i < n; i++) {
            HFr t = i < ell ? fr_sub(Bcoef[i], s.beta_gp_inv) : fr_add(Bcoef[i], s.alpha_gp);
            fr_to_bytes(fD + 32 * (size_t)i, t);
            fr_to_bytes(fBc + 32 * (size_t)i, r[i]);
            fr_to_bytes(fBd + 32 * (size_t)i, fr_mul(zz[i], uvec[i]));
        }
        // keep r_c / r_d until alpha is known: stash them in wG / wGp (weights are initialised next round)
        for (uint32_t i = 0; i < n; i++) { wG[i] = r[i]; wGp[i] = zz[i]; }
        s.tr = tr;
        return;
    }
    const uint32_t R_IPA0 = 5, R_SS = 5 + lg, R_MSM_INIT = 6 + lg, R_MSM0 = 7 + lg, R_FIN = 7 + 2 * lg;
    if (round >= R_IPA0 && round < R_SS) {              // IPA rounds                                   (ipa.py:100-151)
        const uint32_t j = round - R_IPA0;
        if (j == 0) {
            tr.append_point("ipa_step1", outs + 48 * O.C);
            tr.append_point("ipa_step1", outs + 48 * O.D);
            tr.append_fr("ipa_step1", s.z);
            tr.append_point("ipa_step1", outs + 48 * O.Bc);
            tr.append_point("ipa_step1", outs + 48 * O.Bd);
            s.alpha_ipa = tr.challenge("ipa_alpha");
            s.beta_ipa = tr.challenge("ipa_beta");
            for (uint32_t i = 0; i < n; i++) {          // c = r_c + alpha c ; d = r_d + alpha d ; weights reset
                c[i] = fr_add(wG[i], fr_mul(s.alpha_ipa, c[i]));
                d[i] = fr_add(wGp[i], fr_mul(s.alpha_ipa, d[i]));
                wG[i] = fr_one(); wGp[i] = uvec[i];
            }
            s.len = n;
        } else {                                        // absorb the previous round's L/R, fold
            const uint8_t* pr = outs + 48 * (size_t)(O.ipa0 + 4 * (j - 1));
            for (uint32_t k = 0; k < 4; k++) tr.append_point("ipa_loop", pr + 48 * k);
            HFr gam = tr.challenge("ipa_gamma"), gam_inv = fr_inv(gam);
            uint32_t m = s.len / 2;
            for (uint32_t i = 0; i < m; i++) {
                c[i] = fr_add(c[i], fr_mul(gam_inv, c[m + i]));
                d[i] = fr_add(d[i], fr_mul(gam, d[m + i]));
            }
            for (uint32_t L = 0; L < n; L++) if ((L / m) & 1) { wG[L] = fr_mul(wG[L], gam); wGp[L] = fr_mul(wGp[L], gam_inv); }
            s.len = m;
        }
        // outputs of this round: L_C, L_D, R_C, R_D over the leaves
        const uint32_t len = s.len, m = len / 2;
        zero_rows(sh, pb, b, 4, 0);
        uint8_t* fLC = frow(sh, pb, b, 0); uint8_t* fLD = frow(sh, pb, b, 1); uint8_t* fRC = frow(sh, pb, b, 2); uint8_t* fRD = frow(sh, pb, b, 3);
        for (uint32_t L = 0; L < n; L++) {
            uint32_t i = L % m;
            if ((L / m) & 1) {                          // leaf folds into the right half
                fr_to_bytes(fLC + 32 * (size_t)L, fr_mul(c[i], wG[L]));          // MSM(G_R, c_L)
                fr_to_bytes(fRD + 32 * (size_t)L, fr_mul(d[i], wGp[L]));         // MSM(G'_R, d_L)
            } else {
                fr_to_bytes(fRC + 32 * (size_t)L, fr_mul(c[m + i], wG[L]));      // MSM(G_L, c_R)
                fr_to_bytes(fLD + 32 * (size_t)L, fr_mul(d[m + i], wGp[L]));     // MSM(G'_L, d_R)
            }
        }
        fr_to_bytes(fLC + 32 * (size_t)n, fr_mul(s.beta_ipa, ip(c, d + m, m)));  // + H beta <c_L, d_R>
        fr_to_bytes(fRC + 32 * (size_t)n, fr_mul(s.beta_ipa, ip(c + m, d, m)));  // + H beta <c_R, d_L>
        s.tr = tr;
        return;
    }
    if (round == R_SS) {                                // last IPA fold; R, S, cm_T, cm_U, cm_A, cm_B   (curdleproofs.py:92-102, same_scalar.py:39-44)
        {
            const uint8_t* pr = outs + 48 * (size_t)(O.ipa0 + 4 * (lg - 1));
            for (uint32_t k = 0; k < 4; k++) tr.append_point("ipa_loop", pr + 48 * k);
            HFr gam = tr.challenge("ipa_gamma"), gam_inv = fr_inv(gam);
            c[0] = fr_add(c[0], fr_mul(gam_inv, c[1]));                           // c_final, d_final
            d[0] = fr_add(d[0], fr_mul(gam, d[1]));
        }
        s.r_t = prand(pb, sh, b, RO.r_t); s.r_u = prand(pb, sh, b, RO.r_u);
        s.r_a = prand(pb, sh, b, RO.r_a); s.r_b = prand(pb, sh, b, RO.r_b); s.r_k = prand(pb, sh, b, RO.r_k);
        zero_rows(sh, pb, b, 10, 2);
        // order: Rp Sp T1 T2 U1 U2 A1 A2 B1 B2 ; var rows: 0 R' = MSM(vec_R, a), 1 S' = MSM(vec_S, a); the
        // commitments' k R', r_k R', k S', r_k S' are one scalar-mul each of those results (ProveCombine)
        uint8_t *vR = vrow(sh, pb, b, 0), *vS = vrow(sh, pb, b, 1);
        for (uint32_t i = 0; i < ell; i++) {
            fr_to_bytes(vR + 32 * (size_t)i, a[i]);
            cpgh::copy32(vS + 32 * (size_t)i, vR + 32 * (size_t)i);
        }
        const size_t iH = n, iGt = n + 1, iGu = n + 2;
        fr_to_bytes(frow(sh, pb, b, 2) + 32 * iGt, s.r_t);      // cm_T = (G_t r_t, R' k + H r_t)
        fr_to_bytes(frow(sh, pb, b, 3) + 32 * iH, s.r_t);
        fr_to_bytes(frow(sh, pb, b, 4) + 32 * iGu, s.r_u);      // cm_U = (G_u r_u, S' k + H r_u)
        fr_to_bytes(frow(sh, pb, b, 5) + 32 * iH, s.r_u);
        fr_to_bytes(frow(sh, pb, b, 6) + 32 * iGt, s.r_a);      // cm_A = (G_t r_a, R' r_k + H r_a)
        fr_to_bytes(frow(sh, pb, b, 7) + 32 * iH, s.r_a);
        fr_to_bytes(frow(sh, pb, b, 8) + 32 * iGu, s.r_b);      // cm_B = (G_u r_b, S' r_k + H r_b)
        fr_to_bytes(frow(sh, pb, b, 9) + 32 * iH, s.r_b);
        s.tr = tr;
        return;
    }
    if (round == R_MSM_INIT) {                          // SameScalar responses; A', B_a, B_t, B_u       (same_scalar.py:46-63, same_msm.py:73-77)
        const uint32_t ss[10] = {O.Rp, O.Sp, O.T1, O.T2, O.U1, O.U2, O.A1, O.A2, O.B1, O.B2};
        for (uint32_t k = 0; k < 10; k++) tr.append_point("sameexp_points", outs + 48 * (size_t)ss[k]);
        HFr alpha = tr.challenge("same_scalar_alpha");
        s.z_k = fr_add(s.r_k, fr_mul(s.k, alpha));
        s.z_t = fr_add(s.r_a, fr_mul(s.r_t, alpha));
        s.z_u = fr_add(s.r_b, fr_mul(s.r_u, alpha));
        // x_wb = a_perm | a_bl | r_t r_u ; r = msm blinders (kept in w2 until alpha_msm is known)
        for (uint32_t i = 0; i < ell + 2; i++) x[i] = aperm[i];
        x[ell + 2] = s.r_t; x[ell + 3] = s.r_u;
        for (uint32_t i = 0; i < n; i++) w2[i] = prand(pb, sh, b, RO.msm_r + i);
        zero_rows(sh, pb, b, 4, 2);
        // outputs: 0 A' 1 B_a 2 B_t 3 B_u ; var rows: 0 B_t (T) 1 B_u (U)
        uint8_t *fAp = frow(sh, pb, b, 0), *fBa = frow(sh, pb, b, 1), *fBt = frow(sh, pb, b, 2), *fBu = frow(sh, pb, b, 3);
        for (uint32_t i = 0; i < n; i++) fr_to_bytes(fAp + 32 * (size_t)i, Acoef[i]);
        fr_to_bytes(fAp + 32 * (size_t)(n + 1), s.r_t);          // + cm_T.T_1 = G_t r_t
        fr_to_bytes(fAp + 32 * (size_t)(n + 2), s.r_u);          // + cm_U.T_1 = G_u r_u
        for (uint32_t L = 0; L < n; L++) fr_to_bytes(fBa + 32 * (size_t)gwb_index(sh, L), w2[L]);
        uint8_t *vT = vrow(sh, pb, b, 0), *vU = vrow(sh, pb, b, 1);
        for (uint32_t i = 0; i < ell; i++) { fr_to_bytes(vT + 32 * (size_t)i, w2[i]); cpgh::copy32(vU + 32 * (size_t)i, vT + 32 * (size_t)i); }
        fr_to_bytes(fBt + 32 * (size_t)n, w2[ell + 2]);          // T_wb = vec_T | 0 0 H 0
        fr_to_bytes(fBu + 32 * (size_t)n, w2[ell + 3]);          // U_wb = vec_U | 0 0 0 H
        s.tr = tr;
        return;
    }
    if (round >= R_MSM0 && round < R_FIN) {             // SameMSM rounds                                 (same_msm.py:79-131)
        const uint32_t j = round - R_MSM0;
        if (j == 0) {
            tr.append_point("same_msm_step1", outs + 48 * O.Ap);
            tr.append_point("same_msm_step1", outs + 48 * O.T2);
            tr.append_point("same_msm_step1", outs + 48 * O.U2);
            uint8_t INF[48]; memset(INF, 0, 48); INF[0] = 0xc0;
            const uint8_t* Hb = pb.crs48 + 48 * (size_t)n;
            const uint8_t* tu = pb.tu48 + b * (size_t)(2 * ell) * 48;
            for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", tu + 48 * (size_t)i);
            tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
            tr.append_point("same_msm_step1", Hb); tr.append_point("same_msm_step1", INF);
            for (uint32_t i = 0; i < ell; i++) tr.append_point("same_msm_step1", tu + 48 * (size_t)(ell + i));
            tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", INF);
            tr.append_point("same_msm_step1", INF); tr.append_point("same_msm_step1", Hb);
            tr.append_point("same_msm_step1", outs + 48 * O.Ba);
            tr.append_point("same_msm_step1", outs + 48 * O.Bt);
            tr.append_point("same_msm_step1", outs + 48 * O.Bu);
            s.alpha_msm = tr.challenge("same_msm_alpha");
            for (uint32_t i = 0; i < n; i++) { x[i] = fr_add(w2[i], fr_mul(s.alpha_msm, x[i])); }
            for (uint32_t i = 0; i < n; i++) w2[i] = fr_one();
            s.len = n;
        } else {
            const uint8_t* pr = outs + 48 * (size_t)(O.msm0 + 6 * (j - 1));
            for (uint32_t k = 0; k < 6; k++) tr.append_point("same_msm_loop", pr + 48 * k);
            HFr gam = tr.challenge("same_msm_gamma"), gam_inv = fr_inv(gam);
            uint32_t m = s.len / 2;
            for (uint32_t i = 0; i < m; i++) x[i] = fr_add(x[i], fr_mul(gam_inv, x[m + i]));
            for (uint32_t L = 0; L < n; L++) if ((L / m) & 1) w2[L] = fr_mul(w2[L], gam);
            s.len = m;
        }
        const uint32_t len = s.len, m = len / 2;
        zero_rows(sh, pb, b, 6, 4);
        // outputs: 0 L_A 1 L_T 2 L_U 3 R_A 4 R_T 5 R_U ; var rows: 0 L_T (T) 1 L_U (U) 2 R_T (T) 3 R_U (U)
        uint8_t *fLA = frow(sh, pb, b, 0), *fLT = frow(sh, pb, b, 1), *fLU = frow(sh, pb, b, 2), *fRA = frow(sh, pb, b, 3), *fRT = frow(sh, pb, b, 4), *fRU = frow(sh, pb, b, 5);
        uint8_t *vLT = vrow(sh, pb, b, 0), *vLU = vrow(sh, pb, b, 1), *vRT = vrow(sh, pb, b, 2), *vRU = vrow(sh, pb, b, 3);
        for (uint32_t L = 0; L < n; L++) {
            uint32_t i = L % m;
            bool right = ((L / m) & 1) != 0;
            HFr coef = fr_mul(right ? x[i] : x[m + i], w2[L]);   // L_* = MSM(v[m:], x_L), R_* = MSM(v[:m], x_R)
            uint8_t* fA = right ? fLA : fRA;
            fr_to_bytes(fA + 32 * (size_t)gwb_index(sh, L), coef);
            if (L < ell) {
                uint8_t* vT = right ? vLT : vRT;
                fr_to_bytes(vT + 32 * (size_t)L, coef);
                cpgh::copy32((right ? vLU : vRU) + 32 * (size_t)L, vT + 32 * (size_t)L);
            } else if (L == ell + 2) {
                fr_to_bytes((right ? fLT : fRT) + 32 * (size_t)n, coef);          // H inside T_wb
            } else if (L == ell + 3) {
                fr_to_bytes((right ? fLU : fRU) + 32 * (size_t)n, coef);          // H inside U_wb
            }
        }
        s.tr = tr;
        return;
    }
    if (round == R_FIN) {                               // last SameMSM fold + wire assembly             (curdleproofs.py:275-285)
        {
            const uint8_t* pr = outs + 48 * (size_t)(O.msm0 + 6 * (lg - 1));
            for (uint32_t k = 0; k < 6; k++) tr.append_point("same_msm_loop", pr + 48 * k);
            HFr gam = tr.challenge("same_msm_gamma"), gam_inv = fr_inv(gam);
            x[0] = fr_add(x[0], fr_mul(gam_inv, x[1]));                           // x_final
        }
        uint8_t* w = pb.proof + b * (size_t)(1088 + 480 * (size_t)lg);
        auto pt = [&](uint32_t id) { memcpy(w, outs + 48 * (size_t)id, 48); w += 48; };
        auto sc = [&](const HFr& v) { fr_to_bytes(w, v); w += 32; };
        pt(O.A); pt(O.T1); pt(O.T2); pt(O.U1); pt(O.U2); pt(O.Rp); pt(O.Sp);
        pt(O.B); pt(O.C); sc(s.r_p);
        pt(O.Bc); pt(O.Bd);
        for (uint32_t k = 0; k < 4; k++) {               // L_C[], R_C[], L_D[], R_D[]  (ipa.py:260-270)
            const uint32_t sel[4] = {0, 2, 1, 3};        // stored per round as L_C, L_D, R_C, R_D
            for (uint32_t j = 0; j < lg; j++) pt(O.ipa0 + 4 * j + sel[k]);
        }
        sc(c[0]); sc(d[0]);
        pt(O.A1); pt(O.A2); pt(O.B1); pt(O.B2); sc(s.z_k); sc(s.z_t); sc(s.z_u);
        pt(O.Ba); pt(O.Bt); pt(O.Bu);
        for (uint32_t k = 0; k < 6; k++) for (uint32_t j = 0; j < lg; j++) pt(O.msm0 + 6 * j + k);   // L_A L_T L_U R_A R_T R_U
        sc(x[0]);
        s.tr = tr;
        return;
    }
}

struct ProveStep {                // thread = proof
    static constexpr const char* kName = "ProveStep";
    PShape sh; POut O; PBuffers pb; uint32_t round;
    CPG_HD void operator()(uint64_t b) const { prove_step(sh, O, pb, round, (size_t)b); }
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
// (ProveStep: one thread per proof, 32 warps for 4096 proofs) run under the MSM kernels of the other.
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
    uint8_t *d_in48 = nullptr, *d_tu48 = nullptr, *d_k = nullptr, *d_rand = nullptr, *d_outs = nullptr, *d_fs = nullptr, *d_vs = nullptr, *d_proof = nullptr, *d_err = nullptr;
    uint32_t *d_perm = nullptr, *d_off = nullptr; Aff* d_bases = nullptr; PState* d_st = nullptr; HFr* d_vec = nullptr; Jac *d_fix = nullptr, *d_var = nullptr;
    uint8_t *h_tu = nullptr, *h_outs = nullptr, *h_proof = nullptr, *h_err = nullptr;   // pinned: results leave per lane
    uint8_t *h_in = nullptr, *h_rand = nullptr;   // pinned staging of the two large inputs (pageable caller memory copies at a third of the rate)
    uint32_t* d_k12 = nullptr;    // [B][8] GLV halves of k
    Aff* d_tab = nullptr; uint32_t tab_ts = 0;   // multiples 1..tab_ts of every T_i, U_i: [B][2 ell][tab_ts]
    Jac* d_tab_jac = nullptr; Fq* d_tab_pz = nullptr;   // build scratch for TAB_CHUNK bases at a time
#ifndef CPG_HOST_EMU
    cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
#endif
    std::vector<void*> all() { return {d_in48, d_tu48, d_k, d_rand, d_outs, d_fs, d_vs, d_proof, d_err, d_perm, d_off, d_bases, d_st, d_vec, d_fix, d_var, d_tab, d_tab_jac, d_tab_pz, d_k12}; }
    void release() {
        for (void* q : all()) cpg_free(q);
        cpg_host_free(h_tu); cpg_host_free(h_outs); cpg_host_free(h_proof); cpg_host_free(h_err); cpg_host_free(h_in); cpg_host_free(h_rand);
        h_tu = h_outs = h_proof = h_err = h_in = h_rand = nullptr;
        d_in48 = d_tu48 = d_k = d_rand = d_outs = d_fs = d_vs = d_proof = d_err = nullptr; d_perm = d_off = nullptr; d_bases = nullptr; d_st = nullptr; d_vec = nullptr; d_fix = d_var = nullptr; d_tab = nullptr; d_tab_jac = nullptr; d_tab_pz = nullptr; d_k12 = nullptr;
        cap = 0;
    }
    int reserve(const PShape& sh, size_t proof_len, size_t Bn, uint32_t NOUT, uint32_t ts) {
        if (Bn <= cap && ts == tab_ts) return 0;
        release();
        tab_ts = ts;
        d_k12 = (uint32_t*)cpg_malloc(Bn * 32);
        d_tab = (Aff*)cpg_malloc(sizeof(Aff) * (Bn * 2 * sh.ell * (size_t)ts + 1));
        d_tab_jac = (Jac*)cpg_malloc(sizeof(Jac) * (TAB_CHUNK * (size_t)ts + 1));
        d_tab_pz = (Fq*)cpg_malloc(sizeof(Fq) * (TAB_CHUNK * (size_t)ts + 1));
        const size_t ell = sh.ell, n = sh.n;
        d_in48 = (uint8_t*)cpg_malloc(Bn * 2 * ell * 48);     d_tu48 = (uint8_t*)cpg_malloc(Bn * 2 * ell * 48);
        d_k = (uint8_t*)cpg_malloc(Bn * 32);                  d_rand = (uint8_t*)cpg_malloc(Bn * sh.NR * 32);
        d_outs = (uint8_t*)cpg_malloc(Bn * NOUT * 48);        d_fs = (uint8_t*)cpg_malloc(Bn * P_MAX_OUT * sh.NF * 32);
        d_vs = (uint8_t*)cpg_malloc(Bn * P_MAX_VAR * ell * 32); d_proof = (uint8_t*)cpg_malloc(Bn * proof_len);
        d_err = (uint8_t*)cpg_malloc(Bn * 2 * ell);           d_perm = (uint32_t*)cpg_malloc(Bn * ell * 4);
        d_off = (uint32_t*)cpg_malloc(Bn * P_MAX_VAR * 4);
        d_bases = (Aff*)cpg_malloc(sizeof(Aff) * Bn * 4 * ell); d_st = (PState*)cpg_malloc(sizeof(PState) * Bn);
        d_vec = (HFr*)cpg_malloc(sizeof(HFr) * Bn * PV_COUNT * n);
        d_fix = (Jac*)cpg_malloc(sizeof(Jac) * Bn * P_MAX_OUT); d_var = (Jac*)cpg_malloc(sizeof(Jac) * Bn * P_MAX_VAR);
        h_tu = (uint8_t*)cpg_host_alloc(Bn * 2 * ell * 48); h_outs = (uint8_t*)cpg_host_alloc(Bn * NOUT * 48);
        h_proof = (uint8_t*)cpg_host_alloc(Bn * proof_len);   h_err = (uint8_t*)cpg_host_alloc(Bn * 2 * ell);
        h_in = (uint8_t*)cpg_host_alloc(Bn * 2 * ell * 48);   h_rand = (uint8_t*)cpg_host_alloc(Bn * sh.NR * 32);
        for (void* q : all()) if (!q) { release(); return fail("cpg_prove_batch: device allocation failed"); }
        if (!h_tu || !h_outs || !h_proof || !h_err || !h_in || !h_rand) { release(); return fail("cpg_prove_batch: pinned host allocation failed"); }
        cap = Bn;
        return 0;
    }
};
constexpr int P_MAX_LANES = 4;

struct Prover {
    PShape sh;
    size_t proof_len;             // 1088 + 480 lg (without M)
    std::vector<uint8_t> crs48;
    Aff* d_crs = nullptr; uint8_t* d_crs48 = nullptr;
    void* table = nullptr;        // fixed-base table over vec_G | vec_H | H | G_t | G_u
    int var_window = 0;
    int table_window = 6;         // per-base tables of 2^(c-1) multiples for the T / U MSMs (0: bucket method for those too)
    int nlanes = 2, lastK = 1;
    size_t lane_min = 256;        // proofs per lane below which a batch is not split
    size_t lastB = 0;
    ProverLane lanes[P_MAX_LANES];
#ifndef CPG_HOST_EMU
    cudaEvent_t fork = nullptr;
#endif
    void release() { for (ProverLane& L : lanes) L.release(); }
    // contiguous split of B proofs; small batches stay on one lane (nothing to hide behind)
    int split(size_t B, size_t* first, size_t* count) const {
        int k = (B >= lane_min * (size_t)nlanes) ? nlanes : 1;
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
        const size_t nbases = B * 2 * (size_t)ell;
        for (size_t t0 = 0; t0 < nbases; t0 += TAB_CHUNK) {
            size_t cnt = nbases - t0 < TAB_CHUNK ? nbases - t0 : TAB_CHUNK;
            if (int rc = launch_occ(VarTableBuild{p.tab_ts, p.d_bases, 4 * (uint64_t)ell, 2 * (uint64_t)ell, 2 * (uint64_t)ell, t0, p.d_tab_jac, p.d_tab_pz, p.d_tab}, cnt)) return rc;
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
    pb.st = p.d_st; pb.vec = p.d_vec; pb.outs48 = p.d_outs; pb.fs = p.d_fs; pb.vs = p.d_vs; pb.proof = p.d_proof; pb.B = B;

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
    {
        if (int rc = launch<64>(ProveStep{sh, O, pb, r}, B)) return rc;
        if (r == 7 + 2 * lg) return 0;
        RoundPlan pl = plan_for(r);
        // fixed-base part of every output of the round: B*nout MSMs over the CRS table (rows are output-major)
        if (int rc = cpg_g1_msm_fixed_batched(pr.table, p.d_fs, B * pl.nout, 0, p.d_fix)) return rc;
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
                Xyzz* partial = sc.get<Xyzz>(M * rc_.W);
                if (!partial) return fail("cpg_prove_batch: scratch allocation failed");
                if (int rc = launch<128, 3>(VarTableMsmWindow{ell, p.tab_ts, rc_.W, rc_, (uint32_t)M, p.d_tab, p.d_off, (const uint32_t*)p.d_vs, partial}, ((M + 31) / 32) * 32 * rc_.W)) return rc;
                MsmShape hs; memset(&hs, 0, sizeof hs); hs.W = rc_.W; hs.c = c;
                if (int rc = launch_occ(Horner{hs, partial, p.d_var}, M)) return rc;
            } else if (int rc = cpg_g1_msm_batched_off(p.d_bases, p.d_off, p.d_vs, B * pl.nvar, ell, pr.var_window, p.d_var)) return rc;
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
            if (!rc) rc = d2h_async(L.h_proof, L.d_proof, L.B * p.proof_len);
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

void* cpg_prover_create(const uint8_t* crs_bytes, size_t ell, size_t n_blinders, int fixed_window) {
    if (need_init()) return nullptr;
    size_t n = ell + n_blinders;
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    if (((size_t)1 << lg) != n || n_blinders != 4 || ell < 4 || lg > MAX_LG) { fail("cpg_prover_create: need ell + 4 = 2^k, 3 <= k <= 16"); return nullptr; }
    Prover* p = new Prover;
    p->sh.ell = (uint32_t)ell; p->sh.n = (uint32_t)n; p->sh.lg = lg; p->sh.NF = (uint32_t)n + 3;
    p->sh.NR = PRand((uint32_t)n).NR; p->sh.rounds = 8 + 2 * lg;
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
    if (!rc) { p->table = cpg_fixed_table_create(p->d_crs, n + 3, fixed_window > 0 ? fixed_window : 12); if (!p->table) rc = 1; }
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
    for (ProverLane& L : p->lanes) { if (L.stream) cudaStreamDestroy(L.stream); if (L.done) cudaEventDestroy(L.done); }
    if (p->fork) cudaEventDestroy(p->fork);
#endif
    cpg_fixed_table_free(p->table);
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
int cpg_prover_set_lanes(void* handle, int nlanes, size_t min_proofs_per_lane) {
    if (!handle) return fail("cpg_prover_set_lanes: null prover");
    if (nlanes < 1 || nlanes > P_MAX_LANES) return fail("cpg_prover_set_lanes: 1..4 lanes");
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
        if (int rc = L.reserve(sh, p.proof_len, c, O.NOUT, (p.table_window > 0 && ell <= 2048) ? 1u << (p.table_window - 1) : 0)) return rc;
        L.B = c;
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

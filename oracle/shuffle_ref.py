"""ORACLE (test infrastructure, never on the product path).

Per-proof CPU restatement of the reference's shuffle argument, written against the abstract
G1Point/Scalar surface so the same code runs on the oracle arithmetic (golden vectors, CPU
baseline) and on the CUDA drop-in surface (parity tests).  It follows the reference's
operation, RNG-draw and Fiat-Shamir order exactly, so that for one ``random.seed`` the proof
bytes equal those of the unmodified reference (tests/test_oracle_golden.py: against the committed
fixtures the unmodified reference wrote, and re-generated from /root/reference whenever it is mounted).  Cited sources, all under
/root/reference/curdleproofs/curdleproofs/:
  curdleproofs.py:50-160 (prove) :162-248 (verify) :275-298 (wire) :301-321 (shuffle+commit)
  same_perm.py:27-72 / :74-120      grand_prod.py:29-119 / :121-177
  ipa.py:27-48, :75-153 / :155-233  same_scalar.py:24-69 / :71-111   commitment.py:30
  same_msm.py:50-144 / :146-226     msm_accumulator.py:6-12, :32-68  util.py:21-24
"""
import random

from .bls12381_py import R as ORDER
from .merlin_py import Transcript

N_BLINDERS = 4


class VerifyError(AssertionError):
    pass


class ShuffleRef:
    def __init__(self, G1Point, Scalar, transcript_cls=Transcript, rng=random):
        self.P = G1Point
        self.S = Scalar
        self.T = transcript_cls
        self.rng = rng
        self.zero_pt = G1Point.identity()
        self.gen = G1Point()

    # ---- helpers -------------------------------------------------------------------
    def rand(self):  # util.py:21-24
        return self.S.from_le_bytes(self.rng.randint(1, ORDER - 1).to_bytes(32, "little"))

    def rands(self, n):
        return [self.rand() for _ in range(n)]

    def msm(self, bases, scalars):  # msm_accumulator.py:6-12 (naive loop, zip-truncating)
        acc = self.P.identity()
        for b, s in zip(bases, scalars):
            acc = acc + b * s
        return acc

    def ip(self, a, b):
        assert len(a) == len(b)
        acc = self.S(0)
        for x, y in zip(a, b):
            acc = acc + x * y
        return acc

    def inv(self, x):
        y = x.inverse()
        assert y * x == self.S(1)
        return y

    def spow(self, x, e):
        out = self.S(1)
        while e:
            if e & 1:
                out = out * x
            x = x * x
            e >>= 1
        return out

    @staticmethod
    def pb(pt):
        return bytes(pt.to_compressed_bytes())

    @staticmethod
    def fb(s):
        return bytes(s.to_le_bytes())

    def chal(self, tr, label):
        return self.S.from_le_bytes(tr.challenge_int(label).to_bytes(32, "little"))

    # ---- CRS / inputs ---------------------------------------------------------------
    def make_crs(self, ell, n_blinders=N_BLINDERS):  # crs.py:38-66
        pts = [self.gen * self.rand() for _ in range(ell + n_blinders + 3)]
        vec_G, vec_H = pts[:ell], pts[ell:ell + n_blinders]
        g_sum = self.P.identity()
        for g in vec_G:
            g_sum = g_sum + g
        h_sum = self.P.identity()
        for h in vec_H:
            h_sum = h_sum + h
        return dict(vec_G=vec_G, vec_H=vec_H, H=pts[ell + n_blinders], G_t=pts[ell + n_blinders + 1],
                    G_u=pts[ell + n_blinders + 2], G_sum=g_sum, H_sum=h_sum)

    def crs_to_bytes(self, crs):  # crs.py:93-102
        seq = crs["vec_G"] + crs["vec_H"] + [crs["H"], crs["G_t"], crs["G_u"], crs["G_sum"], crs["H_sum"]]
        return b"".join(self.pb(p) for p in seq)

    def crs_from_bytes(self, data, ell, n_blinders=N_BLINDERS):
        pts = [self.P.from_compressed_bytes_unchecked(data[48 * i:48 * i + 48]) for i in range(ell + n_blinders + 5)]
        return dict(vec_G=pts[:ell], vec_H=pts[ell:ell + n_blinders], H=pts[ell + n_blinders],
                    G_t=pts[ell + n_blinders + 1], G_u=pts[ell + n_blinders + 2],
                    G_sum=pts[ell + n_blinders + 3], H_sum=pts[ell + n_blinders + 4])

    def shuffle_and_commit(self, crs, vec_R, vec_S, perm, k):  # curdleproofs.py:301-321
        ell = len(crs["vec_G"])
        vec_T = [vec_R[i] * k for i in range(len(vec_R))]
        vec_U = [vec_S[i] * k for i in range(len(vec_S))]
        vec_T = [vec_T[i] for i in perm]
        vec_U = [vec_U[i] for i in perm]
        sigma = [self.S(int(i)) for i in perm][:ell]
        m_bl = self.rands(N_BLINDERS)
        M = self.msm(crs["vec_G"], sigma) + self.msm(crs["vec_H"], m_bl)
        return vec_T, vec_U, M, m_bl

    # ---- prover ---------------------------------------------------------------------
    def prove(self, crs, vec_R, vec_S, vec_T, vec_U, M, perm, k, m_bl):
        """Returns the proof as wire bytes (curdleproofs.py:275-285 layout)."""
        S = self.S
        ell = len(vec_R)
        tr = self.T(b"curdleproofs")
        tr.append_all(b"curdleproofs_step1", [self.pb(p) for p in vec_R + vec_S + vec_T + vec_U])
        tr.append(b"curdleproofs_step1", self.pb(M))
        a = [self.chal(tr, b"curdleproofs_vec_a") for _ in range(ell)]

        a_bl = self.rands(N_BLINDERS - 2)
        r_a_prime = a_bl + [S(0), S(0)]
        a_perm = [a[i] for i in perm]
        A = self.msm(crs["vec_G"], a_perm) + self.msm(crs["vec_H"], r_a_prime)

        same_perm = self._prove_same_perm(crs, A, M, a, perm, r_a_prime, m_bl, tr)

        r_t, r_u = self.rand(), self.rand()
        Rp = self.msm(vec_R, a)
        Sp = self.msm(vec_S, a)
        cm_T = (crs["G_t"] * r_t, Rp * k + crs["H"] * r_t)
        cm_U = (crs["G_u"] * r_u, Sp * k + crs["H"] * r_u)

        same_scalar = self._prove_same_scalar(crs, Rp, Sp, cm_T, cm_U, k, r_t, r_u, tr)

        A_prime = A + cm_T[0] + cm_U[0]
        Z = self.zero_pt
        G_wb = crs["vec_G"] + crs["vec_H"][:N_BLINDERS - 2] + [crs["G_t"], crs["G_u"]]
        T_wb = vec_T + [Z, Z, crs["H"], Z]
        U_wb = vec_U + [Z, Z, Z, crs["H"]]
        x_wb = a_perm + a_bl + [r_t, r_u]
        same_msm = self._prove_same_msm(G_wb, A_prime, cm_T[1], cm_U[1], T_wb, U_wb, x_wb, tr)

        return b"".join([self.pb(A), self.pb(cm_T[0]), self.pb(cm_T[1]), self.pb(cm_U[0]), self.pb(cm_U[1]),
                         self.pb(Rp), self.pb(Sp), same_perm, same_scalar, same_msm])

    def _prove_same_perm(self, crs, A, M, a, perm, a_bl, m_bl, tr):  # same_perm.py:27-72
        S = self.S
        vec_G, vec_H = crs["vec_G"], crs["vec_H"]
        ell = len(vec_G)
        tr.append_all(b"same_perm_step1", [self.pb(A), self.pb(M)])
        tr.append_all(b"same_perm_step1", [self.fb(x) for x in a])
        alpha = self.chal(tr, b"same_perm_alpha")
        beta = self.chal(tr, b"same_perm_beta")
        a_perm = [a[i] for i in perm]
        factors = [ai + S(int(m)) * alpha + beta for ai, m in zip(a_perm, perm)]
        gprod = S(1)
        for f in factors:
            gprod = gprod * f
        B = (A + M * alpha) + self.msm(vec_G, [beta] * ell)
        b_bl = [a_bl[i] + alpha * m_bl[i] for i in range(len(a_bl))]
        gp = self._prove_gprod(vec_G, vec_H, crs["H"], B, gprod, factors, b_bl, tr)
        return self.pb(B) + gp

    def _prove_gprod(self, vec_G, vec_H, U, B, gprod, b, b_bl, tr):  # grand_prod.py:29-119
        S = self.S
        nb = len(b_bl)
        ell = len(vec_G)
        tr.append(b"gprod_step1", self.pb(B))
        tr.append(b"gprod_step1", self.fb(gprod))
        alpha = self.chal(tr, b"gprod_alpha")
        c = [S(1)]
        for i in range(ell - 1):
            c.append(c[i] * b[i])
        c_bl = self.rands(nb)
        C = self.msm(vec_G, c) + self.msm(vec_H, c_bl)
        rb_alpha = [x + alpha for x in b_bl]
        r_p = self.ip(rb_alpha, c_bl)
        tr.append(b"gprod_step2", self.pb(C))
        tr.append(b"gprod_step2", self.fb(r_p))
        beta = self.chal(tr, b"gprod_beta")
        beta_inv = self.inv(beta)

        G_prime = []
        pw = beta_inv
        for g in vec_G:
            G_prime.append(g * pw)
            pw = pw * beta_inv
        beta_inv_l1 = self.spow(beta_inv, ell + 1)
        H_prime = [h * beta_inv_l1 for h in vec_H]

        b_prime = []
        pw = beta
        for bi in b:
            b_prime.append(bi * pw)
            pw = pw * beta
        d = []
        beta_pows = []
        pw = S(1)
        for bp in b_prime:
            d.append(bp - pw)
            beta_pows.append(pw)
            pw = pw * beta
        beta_l1 = self.spow(beta, ell + 1)
        d_bl = [beta_l1 * x for x in rb_alpha]
        alphabeta = [alpha * beta_l1 for _ in range(nb)]
        D = B - self.msm(G_prime, beta_pows) + self.msm(H_prime, alphabeta)

        G_all = vec_G + vec_H
        Gp_all = G_prime + H_prime
        z = r_p * beta_l1 + gprod * self.spow(beta, ell) - S(1)
        c = c + c_bl
        d = d + d_bl
        # prover self-checks of the reference (grand_prod.py:103-105)
        assert self.ip(c, d) == z
        assert self.msm(G_all, c) == C
        assert self.msm(Gp_all, d) == D
        ipa = self._prove_ipa(G_all, Gp_all, U, C, D, z, c, d, tr)
        return self.pb(C) + self.fb(r_p) + ipa

    def _ipa_blinders(self, c, d):  # ipa.py:27-48
        n = len(c)
        r = self.rands(n)
        z = self.rands(n - 2)
        omega = self.ip(r, d) + self.ip(z[:n - 2], c[:n - 2])
        delta = self.ip(r[:n - 2], z[:n - 2])
        inv_c = self.inv(c[n - 2])
        last_z = (r[n - 2] * inv_c * omega - delta) * self.inv(-r[n - 2] * inv_c * c[n - 1] + r[n - 1])
        pen_z = -inv_c * (last_z * c[n - 1] + omega)
        z = z + [pen_z, last_z]
        assert self.ip(r, d) + self.ip(z, c) == self.S(0)
        assert self.ip(r, z) == self.S(0)
        return r, z

    def _prove_ipa(self, G, Gp, U, C, D, z, c, d, tr):  # ipa.py:75-153
        n = len(c)
        assert n & (n - 1) == 0 and n == len(d)
        r_c, r_d = self._ipa_blinders(c, d)
        B_c = self.msm(G, r_c)
        B_d = self.msm(Gp, r_d)
        tr.append_all(b"ipa_step1", [self.pb(C), self.pb(D)])
        tr.append(b"ipa_step1", self.fb(z))
        tr.append_all(b"ipa_step1", [self.pb(B_c), self.pb(B_d)])
        alpha = self.chal(tr, b"ipa_alpha")
        beta = self.chal(tr, b"ipa_beta")
        c = [r_c[i] + alpha * c[i] for i in range(n)]
        d = [r_d[i] + alpha * d[i] for i in range(n)]
        H = U * beta
        LC, RC, LD, RD = [], [], [], []
        G = list(G)
        Gp = list(Gp)
        while len(c) > 1:
            n //= 2
            cL, cR, dL, dR = c[:n], c[n:], d[:n], d[n:]
            GL, GR, GpL, GpR = G[:n], G[n:], Gp[:n], Gp[n:]
            L_C = self.msm(GR, cL) + H * self.ip(cL, dR)
            L_D = self.msm(GpL, dR)
            R_C = self.msm(GL, cR) + H * self.ip(cR, dL)
            R_D = self.msm(GpR, dL)
            LC.append(L_C); RC.append(R_C); LD.append(L_D); RD.append(R_D)
            tr.append_all(b"ipa_loop", [self.pb(L_C), self.pb(L_D), self.pb(R_C), self.pb(R_D)])
            gamma = self.chal(tr, b"ipa_gamma")
            gamma_inv = self.inv(gamma)
            c = [cL[i] + gamma_inv * cR[i] for i in range(n)]
            d = [dL[i] + gamma * dR[i] for i in range(n)]
            G = [GL[i] + GR[i] * gamma for i in range(n)]
            Gp = [GpL[i] + GpR[i] * gamma_inv for i in range(n)]
        out = [self.pb(B_c), self.pb(B_d)]
        for vec in (LC, RC, LD, RD):
            out += [self.pb(p) for p in vec]
        return b"".join(out) + self.fb(c[0]) + self.fb(d[0])

    def _prove_same_scalar(self, crs, Rp, Sp, cm_T, cm_U, k, r_t, r_u, tr):  # same_scalar.py:24-69
        r_a, r_b, r_k = self.rand(), self.rand(), self.rand()
        cm_A = (crs["G_t"] * r_a, Rp * r_k + crs["H"] * r_a)
        cm_B = (crs["G_u"] * r_b, Sp * r_k + crs["H"] * r_b)
        tr.append_all(b"sameexp_points", [self.pb(p) for p in
                                          (Rp, Sp, cm_T[0], cm_T[1], cm_U[0], cm_U[1], cm_A[0], cm_A[1], cm_B[0], cm_B[1])])
        alpha = self.chal(tr, b"same_scalar_alpha")
        z_k = r_k + k * alpha
        z_t = r_a + r_t * alpha
        z_u = r_b + r_u * alpha
        return b"".join([self.pb(cm_A[0]), self.pb(cm_A[1]), self.pb(cm_B[0]), self.pb(cm_B[1]),
                         self.fb(z_k), self.fb(z_t), self.fb(z_u)])

    def _prove_same_msm(self, G, A, Z_t, Z_u, T, U, x, tr):  # same_msm.py:50-144
        n = len(x)
        assert n & (n - 1) == 0
        r = self.rands(n)
        B_a = self.msm(G, r)
        B_t = self.msm(T, r)
        B_u = self.msm(U, r)
        tr.append_all(b"same_msm_step1", [self.pb(A), self.pb(Z_t), self.pb(Z_u)])
        tr.append_all(b"same_msm_step1", [self.pb(p) for p in T + U])
        tr.append_all(b"same_msm_step1", [self.pb(B_a), self.pb(B_t), self.pb(B_u)])
        alpha = self.chal(tr, b"same_msm_alpha")
        x = [r[i] + alpha * x[i] for i in range(n)]
        LA, LT, LU, RA, RT, RU = [], [], [], [], [], []
        G, T, U = list(G), list(T), list(U)
        while len(x) > 1:
            n //= 2
            xL, xR = x[:n], x[n:]
            pts = [self.msm(G[n:], xL), self.msm(T[n:], xL), self.msm(U[n:], xL),
                   self.msm(G[:n], xR), self.msm(T[:n], xR), self.msm(U[:n], xR)]
            for lst, p in zip((LA, LT, LU, RA, RT, RU), pts):
                lst.append(p)
            tr.append_all(b"same_msm_loop", [self.pb(p) for p in pts])
            gamma = self.chal(tr, b"same_msm_gamma")
            gamma_inv = self.inv(gamma)
            x = [xL[i] + gamma_inv * xR[i] for i in range(n)]
            T = [T[i] + T[n + i] * gamma for i in range(n)]
            U = [U[i] + U[n + i] * gamma for i in range(n)]
            G = [G[i] + G[n + i] * gamma for i in range(n)]
        out = [self.pb(B_a), self.pb(B_t), self.pb(B_u)]
        for vec in (LA, LT, LU, RA, RT, RU):
            out += [self.pb(p) for p in vec]
        return b"".join(out) + self.fb(x[0])

    # ---- wire format ------------------------------------------------------------------
    def parse_proof(self, data, n):
        """Split proof wire bytes into named points/scalars (curdleproofs.py:287-298 and nested)."""
        lg = n.bit_length() - 1
        if 1 << lg != n:
            raise Exception("n not a power of 2")
        if len(data) < 1088 + 480 * lg:
            raise Exception("proof too short")
        pos = [0]

        def g1():
            p = self.P.from_compressed_bytes_unchecked(data[pos[0]:pos[0] + 48])
            pos[0] += 48
            return p

        def fr():
            s = self.S.from_le_bytes(data[pos[0]:pos[0] + 32])
            pos[0] += 32
            return s

        pr = {}
        pr["A"] = g1(); pr["cm_T"] = (g1(), g1()); pr["cm_U"] = (g1(), g1()); pr["R"] = g1(); pr["S"] = g1()
        pr["B"] = g1(); pr["C"] = g1(); pr["r_p"] = fr()
        pr["B_c"] = g1(); pr["B_d"] = g1()
        for name in ("L_C", "R_C", "L_D", "R_D"):
            pr[name] = [g1() for _ in range(lg)]
        pr["c_final"] = fr(); pr["d_final"] = fr()
        pr["cm_A"] = (g1(), g1()); pr["cm_B"] = (g1(), g1())
        pr["z_k"] = fr(); pr["z_t"] = fr(); pr["z_u"] = fr()
        pr["B_a"] = g1(); pr["B_t"] = g1(); pr["B_u"] = g1()
        for name in ("L_A", "L_T", "L_U", "R_A", "R_T", "R_U"):
            pr[name] = [g1() for _ in range(lg)]
        pr["x_final"] = fr()
        return pr

    # ---- verifier ---------------------------------------------------------------------
    def verify(self, crs, vec_R, vec_S, vec_T, vec_U, M, proof_bytes):
        """Raises (AssertionError / Exception / ValueError) exactly where the reference does."""
        S = self.S
        vec_G, vec_H = crs["vec_G"], crs["vec_H"]
        ell = len(vec_R)
        n = len(vec_G) + len(vec_H)
        pr = self.parse_proof(proof_bytes, n)
        tr = self.T(b"curdleproofs")
        acc = _Accumulator(self)
        if vec_T[0] == self.zero_pt:
            raise Exception("vec_T[0] is infinity")
        tr.append_all(b"curdleproofs_step1", [self.pb(p) for p in vec_R + vec_S + vec_T + vec_U])
        tr.append(b"curdleproofs_step1", self.pb(M))
        a = [self.chal(tr, b"curdleproofs_vec_a") for _ in range(ell)]

        # same permutation (same_perm.py:74-120)
        A = pr["A"]
        tr.append_all(b"same_perm_step1", [self.pb(A), self.pb(M)])
        tr.append_all(b"same_perm_step1", [self.fb(x) for x in a])
        alpha = self.chal(tr, b"same_perm_alpha")
        beta = self.chal(tr, b"same_perm_beta")
        gprod = S(1)
        for i in range(len(vec_G)):
            if i < len(a):
                gprod = gprod * (a[i] + S(i) * alpha + beta)
        acc.check((pr["B"] - A) - (M * alpha), vec_G, [beta] * len(vec_G))

        # grand product (grand_prod.py:121-177)
        B = pr["B"]
        ellg = len(vec_G)
        tr.append(b"gprod_step1", self.pb(B))
        tr.append(b"gprod_step1", self.fb(gprod))
        alpha = self.chal(tr, b"gprod_alpha")
        tr.append(b"gprod_step2", self.pb(pr["C"]))
        tr.append(b"gprod_step2", self.fb(pr["r_p"]))
        beta = self.chal(tr, b"gprod_beta")
        beta_inv = self.inv(beta)
        u = []
        pw = beta_inv
        for _ in range(ellg):
            u.append(pw)
            pw = pw * beta_inv
        u += [self.spow(beta_inv, ellg + 1)] * N_BLINDERS
        D = B - crs["G_sum"] * beta_inv + crs["H_sum"] * alpha
        z = pr["r_p"] * self.spow(beta, ellg + 1) + gprod * self.spow(beta, ellg) - S(1)

        # inner product argument (ipa.py:188-233)
        G_all = vec_G + vec_H
        nn = len(G_all)
        tr.append_all(b"ipa_step1", [self.pb(pr["C"]), self.pb(D)])
        tr.append(b"ipa_step1", self.fb(z))
        tr.append_all(b"ipa_step1", [self.pb(pr["B_c"]), self.pb(pr["B_d"])])
        alpha = self.chal(tr, b"ipa_alpha")
        beta = self.chal(tr, b"ipa_beta")
        gam, gam_inv, s = self._round_scalars(tr, b"ipa_loop", b"ipa_gamma", nn,
                                              [pr["L_C"], pr["L_D"], pr["R_C"], pr["R_D"]])
        s_inv = [self.inv(x) for x in s]
        rhs = [pr["c_final"] * x for x in s] + [pr["c_final"] * pr["d_final"] * beta]
        H = crs["H"] * beta
        C_a = pr["B_c"] + pr["C"] * alpha + H * (alpha * alpha * z)
        lhs = self.msm(pr["L_C"], gam) + C_a + self.msm(pr["R_C"], gam_inv)
        acc.check(lhs, G_all + [crs["H"]], rhs)
        d_div_s = [pr["d_final"] * (si * ui) for si, ui in zip(s_inv, u)]
        D_a = pr["B_d"] + D * alpha
        lhs = self.msm(pr["L_D"], gam) + D_a + self.msm(pr["R_D"], gam_inv)
        acc.check(lhs, G_all, d_div_s)

        # same scalar (same_scalar.py:71-111): exact equalities, not accumulated
        cm_T, cm_U, cm_A, cm_B = pr["cm_T"], pr["cm_U"], pr["cm_A"], pr["cm_B"]
        tr.append_all(b"sameexp_points", [self.pb(p) for p in
                                          (pr["R"], pr["S"], cm_T[0], cm_T[1], cm_U[0], cm_U[1],
                                           cm_A[0], cm_A[1], cm_B[0], cm_B[1])])
        alpha = self.chal(tr, b"same_scalar_alpha")
        e1 = (crs["G_t"] * pr["z_t"], pr["R"] * pr["z_k"] + crs["H"] * pr["z_t"])
        e2 = (crs["G_u"] * pr["z_u"], pr["S"] * pr["z_k"] + crs["H"] * pr["z_u"])
        c1 = (cm_A[0] + cm_T[0] * alpha, cm_A[1] + cm_T[1] * alpha)
        c2 = (cm_B[0] + cm_U[0] * alpha, cm_B[1] + cm_U[1] * alpha)
        if not (e1[0] == c1[0] and e1[1] == c1[1] and e2[0] == c2[0] and e2[1] == c2[1]):
            raise VerifyError("same-scalar check failed")

        # same MSM (same_msm.py:184-226)
        A_prime = A + cm_T[0] + cm_U[0]
        Zp = self.zero_pt
        G_wb = vec_G + vec_H[:N_BLINDERS - 2] + [crs["G_t"], crs["G_u"]]
        T_wb = vec_T + [Zp, Zp, crs["H"], Zp]
        U_wb = vec_U + [Zp, Zp, Zp, crs["H"]]
        nm = len(T_wb)
        tr.append_all(b"same_msm_step1", [self.pb(A_prime), self.pb(cm_T[1]), self.pb(cm_U[1])])
        tr.append_all(b"same_msm_step1", [self.pb(p) for p in T_wb + U_wb])
        tr.append_all(b"same_msm_step1", [self.pb(pr["B_a"]), self.pb(pr["B_t"]), self.pb(pr["B_u"])])
        alpha = self.chal(tr, b"same_msm_alpha")
        gam, gam_inv, s = self._round_scalars(tr, b"same_msm_loop", b"same_msm_gamma", nm,
                                              [pr["L_A"], pr["L_T"], pr["L_U"], pr["R_A"], pr["R_T"], pr["R_U"]])
        xs = [pr["x_final"] * si for si in s]
        for Bx, Zx, Lx, Rx, bases in ((pr["B_a"], A_prime, pr["L_A"], pr["R_A"], G_wb),
                                      (pr["B_t"], cm_T[1], pr["L_T"], pr["R_T"], T_wb),
                                      (pr["B_u"], cm_U[1], pr["L_U"], pr["R_U"], U_wb)):
            lhs = self.msm(Lx, gam) + (Bx + Zx * alpha) + self.msm(Rx, gam_inv)
            acc.check(lhs, bases, xs)

        acc.check(pr["R"], vec_R, a)
        acc.check(pr["S"], vec_S, a)
        acc.finish()

    def _round_scalars(self, tr, loop_label, gamma_label, n, vecs):
        """Per-round challenges and the s-vector (ipa.py:155-186 / same_msm.py:146-182):
        challenge j multiplies s_i when bit j of i, counted from the top, is set (util.py:71-78)."""
        lg = len(vecs[0])
        if lg >= 32:
            raise Exception("too many rounds")
        if 1 << lg != n:
            raise Exception("n != 2**lg")
        gam = []
        for j in range(lg):
            tr.append_all(loop_label, [self.pb(v[j]) for v in vecs])
            gam.append(self.chal(tr, gamma_label))
        gam_inv = [self.inv(g) for g in gam]
        s = []
        for i in range(n):
            acc = self.S(1)
            for j in range(lg):
                if (i >> (lg - 1 - j)) & 1:
                    acc = acc * gam[j]
            s.append(acc)
        return gam, gam_inv, s

    def is_valid(self, crs, vec_R, vec_S, vec_T, vec_U, M, proof_bytes):
        """bool wrapper in the style of IsValidWhiskShuffleProof (whisk_interface.py:74-87)."""
        try:
            self.verify(crs, vec_R, vec_S, vec_T, vec_U, M, proof_bytes)
            return True
        except Exception:
            return False


class _Accumulator:
    """msm_accumulator.py:32-68: random linear combination of MSM checks, keyed by compressed base."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.A_c = ctx.P.identity()
        self.table = {}

    def check(self, C, bases, scalars):
        ctx = self.ctx
        rho = ctx.rand()
        self.A_c = self.A_c + C * rho
        for b, s in zip(bases, scalars):
            if b == ctx.zero_pt:
                continue
            key = ctx.pb(b)
            prev = self.table.get(key)
            term = rho * s
            self.table[key] = term if prev is None else prev + term
        return None

    def finish(self):
        ctx = self.ctx
        keys = list(self.table.keys())
        bases = [ctx.P.from_compressed_bytes_unchecked(k) for k in keys]
        got = ctx.msm(bases, [self.table[k] for k in keys])
        if not got == self.A_c:
            raise VerifyError("accumulated MSM check failed")

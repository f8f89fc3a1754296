"""NumPy port of the B200 solver's algorithm.  TEST INFRASTRUCTURE ONLY.

This is NOT the reference's algorithm (that is oracle/lqp.py + oracle/osqp_port.py); it is a
CPU restatement of what `pmpc_b200/csrc/*.cu` computes — structure-exploiting ADMM whose
z-update is a per-particle Riccati recursion, particles coupled through a Schur complement on
the shared controls, plus an active-set polish — used to debug the kernels stage by stage and
as the "port" CPU baseline of the same arithmetic.  Vectorised over particles; the stage loop
is sequential exactly like the device code.

Problem, 0-based stages j = 0..N-1, per particle i:
    x_j = A_j x_{j-1} + B_j u_j + c_j            (A_0 := 0; c_j = f_j - B_j ubar_j - A_j xbar_{j-1})
    J_i = sum_j 1/2 x_j'Qs_j x_j + q_j'x_j + 1/2 u_j'Rs_j u_j + r_j'u_j + 1/2 s sum_{j>=1} |u_j-u_{j-1}|^2
    u_{i,j} = uc_j for j < Nc;  lu <= u <= uu;  lx <= x <= ux.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _sym_upper(Mx):
    """OSQP consumes the upper triangle of P: mirror it."""
    U = np.triu(Mx)
    return U + np.swapaxes(np.triu(Mx, 1), -1, -2)


@dataclass
class Settings:
    rho: float = 1.0
    alpha: float = 1.6
    eps_abs: float = 1e-9
    eps_rel: float = 1e-9
    max_iter: int = 4000
    check_every: int = 10
    adapt_rho: bool = True
    adapt_tol: float = 3.0
    polish: bool = True
    polish_iters: int = 6
    kkt_tol: float = 1e-9


@dataclass
class State:
    """ADMM warm-start state carried between SCP iterations."""

    w_u: np.ndarray = None  # (M,N,m) (shared stages: row of particle 0 is authoritative)
    y_u: np.ndarray = None
    w_x: np.ndarray = None
    y_x: np.ndarray = None
    rho: float = None


class RiccatiADMM:
    def __init__(self, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx=None, ux=None, lu=None,
                 uu=None, reg_x=0.0, reg_u=0.0, slew_reg=None, slew_reg0=None, slew_um1=None, Nc=-1):
        f = np.asarray(f, float)
        M, N, n = f.shape
        m = np.asarray(fu).shape[-1]
        self.M, self.N, self.n, self.m = M, N, n, m
        Nc = N if Nc < 0 else int(Nc)
        self.Nc = Nc
        nan_absent = lambda *a: any(z is None for z in a) or any(np.isnan(np.asarray(z)).any() for z in a)
        s = np.zeros(M) if nan_absent(slew_reg) else np.broadcast_to(np.asarray(slew_reg, float), (M,)).copy()
        if nan_absent(slew_reg0, slew_um1):
            s0, um1 = np.zeros(M), np.zeros((M, m))
        else:
            s0 = np.broadcast_to(np.asarray(slew_reg0, float), (M,)).copy()
            um1 = np.asarray(slew_um1, float).reshape(M, m)
        self.s = s
        self.has_ubox = not nan_absent(lu, uu)
        self.has_xbox = not nan_absent(lx, ux)
        fx = np.asarray(fx, float).reshape(M, N, n, n)
        fu = np.asarray(fu, float).reshape(M, N, n, m)
        Q = np.asarray(Q, float).reshape(M, N, n, n)
        R = np.asarray(R, float).reshape(M, N, m, m)
        X_prev, U_prev = np.asarray(X_prev, float).reshape(M, N, n), np.asarray(U_prev, float).reshape(M, N, m)
        X_ref, U_ref = np.asarray(X_ref, float).reshape(M, N, n), np.asarray(U_ref, float).reshape(M, N, m)
        reg_x = np.broadcast_to(np.asarray(reg_x, float), (M,))
        reg_u = np.broadcast_to(np.asarray(reg_u, float), (M,))
        self.A = fx.copy()
        self.A[:, 0] = 0.0
        self.B = fu
        c = f - np.einsum("ijrt,ijt->ijr", fu, U_prev)
        c[:, 1:] -= np.einsum("ijrt,ijt->ijr", fx[:, 1:], X_prev[:, :-1])
        self.c = c
        self.Qs = _sym_upper(Q) + reg_x[:, None, None, None] * np.eye(n)
        self.Rs = _sym_upper(R) + reg_u[:, None, None, None] * np.eye(m)
        self.Rs[:, 0] += s0[:, None, None] * np.eye(m)
        if N == 1:
            self.Rs[:, 0] += s[:, None, None] * np.eye(m)  # lqp_utils.jl:31-37 quirk
        self.q = -np.einsum("ijrt,ijt->ijr", Q, X_ref) - reg_x[:, None, None] * X_prev
        self.r = -np.einsum("ijrt,ijt->ijr", R, U_ref) - reg_u[:, None, None] * U_prev
        if Nc >= 1:
            self.r[:, 0] -= s0[:, None] * um1  # dropped when Nc == 0 (reference quirk, App. B.6)
        inf = np.inf
        self.lu = np.asarray(lu, float).reshape(M, N, m).copy() if self.has_ubox else np.full((M, N, m), -inf)
        self.uu = np.asarray(uu, float).reshape(M, N, m).copy() if self.has_ubox else np.full((M, N, m), inf)
        self.lx = np.asarray(lx, float).reshape(M, N, n) if self.has_xbox else np.full((M, N, n), -inf)
        self.ux = np.asarray(ux, float).reshape(M, N, n) if self.has_xbox else np.full((M, N, n), inf)
        # shared block: particle 1's bounds (lqp_utils.jl:329-330)
        self.lu[:, :Nc] = self.lu[0:1, :Nc]
        self.uu[:, :Nc] = self.uu[0:1, :Nc]
        self.x0 = np.asarray(x0, float).reshape(M, n)
        self.stats = dict(admm_iters=0, factorizations=0, polish_attempts=0, polish_ok=False)

    # ------------------------------------------------------------------ factor (K2)
    def factor(self, rho_u, rho_x, act_u=None, ):
        """Backward matrix recursion over the free stages.  `act_u` (M,N,m) int8 marks control
        components fixed at a bound (polish); rho_* are the ADMM penalties (0 in polish)."""
        M, N, n, m, Nc, s = self.M, self.N, self.n, self.m, self.Nc, self.s
        K = np.zeros((M, N, m, n))
        Hinv = np.zeros((M, N, m, m))  # masked inverse E
        Hfull = np.zeros((M, N, m, m))
        SBW = np.zeros((M, N, n, m))  # S B + W (needed for u_fix terms)
        P = np.zeros((M, n, n))
        Pst = np.zeros((M, N, n, n))  # P_j (value Hessian *after* stage j), for vector passes: S_j = Qh_j + P_j
        W = np.zeros((M, n, m))
        Z = np.zeros((M, m, m))
        Wst = np.zeros((M, N, n, m))
        Zst = np.zeros((M, N, m, m))
        I_m = np.eye(m)
        for j in range(N - 1, Nc - 1, -1):
            Pst[:, j], Wst[:, j], Zst[:, j] = P, W, Z
            S = self.Qs[:, j] + rho_x * np.eye(n) + P
            B, A = self.B[:, j], self.A[:, j]
            SB = S @ B + W
            H = self.Rs[:, j] + rho_u * I_m + (s[:, None, None] * I_m if j >= 1 else 0.0)
            H = H + np.swapaxes(B, -1, -2) @ (S @ B) + np.swapaxes(B, -1, -2) @ W + np.swapaxes(W, -1, -2) @ B + Z
            G = np.swapaxes(SB, -1, -2) @ A
            Hfull[:, j], SBW[:, j] = H, SB
            Hm = H.copy()
            if act_u is not None:
                a = act_u[:, j] != 0  # (M,m)
                Hm = np.where(a[:, :, None] | a[:, None, :], 0.0, Hm) + a[:, :, None] * I_m
            E = np.linalg.inv(Hm)
            if act_u is not None:
                E = np.where(a[:, :, None] | a[:, None, :], 0.0, E)
            Kj = -E @ G
            K[:, j], Hinv[:, j] = Kj, E
            P = np.swapaxes(A, -1, -2) @ (S @ A) + np.swapaxes(G, -1, -2) @ Kj
            P = 0.5 * (P + np.swapaxes(P, -1, -2))
            W = -s[:, None, None] * np.swapaxes(Kj, -1, -2)
            Z = s[:, None, None] * I_m - (s**2)[:, None, None] * E
        self.K, self.Hinv, self.Hfull, self.SBW = K, Hinv, Hfull, SBW
        self.Pst, self.Wst, self.Zst = Pst, Wst, Zst
        self.P_c, self.W_c, self.Z_c = P, W, Z  # value function at the consensus boundary
        self.rho_u, self.rho_x = rho_u, rho_x
        self.stats["factorizations"] += 1
        # shared-block Hessian by homogeneous adjoint passes (K3)
        if Nc >= 1:
            D = Nc * m
            Hc = np.zeros((D, D))
            for k in range(D):
                e = np.zeros((M, Nc, m))
                e[:, k // m, k % m] = 1.0
                Hc[:, k] = self._shared_grad(e, None, None, None, homogeneous=True).sum(0).reshape(-1)
            self.Hc_sum = 0.5 * (Hc + Hc.T)

    # ------------------------------------------------------------------ shared-block gradient
    def _shared_grad(self, uc, qh, p_c, z_c, homogeneous=False):
        """d phi_i / d uc for every particle: forward rollout over the Nc shared stages then
        adjoint sweep.  Returns (M,Nc,m)."""
        M, N, n, m, Nc, s = self.M, self.N, self.n, self.m, self.Nc, self.s
        X = np.zeros((M, Nc, n))
        x = np.zeros((M, n))
        for j in range(Nc):
            x = np.einsum("irt,it->ir", self.A[:, j], x) + np.einsum("irt,it->ir", self.B[:, j], uc[:, j])
            if not homogeneous:
                x = x + self.c[:, j]
            X[:, j] = x
        g = np.zeros((M, Nc, m))
        a = np.zeros((M, n))
        for j in range(Nc - 1, -1, -1):
            Qh = self.Qs[:, j] + self.rho_x * np.eye(n)
            a_new = np.einsum("irt,it->ir", Qh, X[:, j])
            if not homogeneous:
                a_new = a_new + qh[:, j]
            gu = np.einsum("irt,it->ir", self.Rs[:, j], uc[:, j])
            if not homogeneous:
                gu = gu + self.r[:, j]
            if j == Nc - 1:
                a_new = a_new + np.einsum("irt,it->ir", self.P_c, X[:, j]) + np.einsum("irt,it->ir", self.W_c, uc[:, j])
                gu = gu + np.einsum("itr,it->ir", self.W_c, X[:, j]) + np.einsum("irt,it->ir", self.Z_c, uc[:, j])
                if not homogeneous:
                    a_new = a_new + p_c
                    gu = gu + z_c
            else:
                a_new = a_new + np.einsum("itr,it->ir", self.A[:, j + 1], a)
                gu = gu + s[:, None] * (uc[:, j] - uc[:, j + 1])
            if j >= 1:
                gu = gu + s[:, None] * (uc[:, j] - uc[:, j - 1])
            a = a_new
            g[:, j] = gu + np.einsum("itr,it->ir", self.B[:, j], a)
        return g

    # ------------------------------------------------------------------ solve (K4)
    def solve_lq(self, rh, qh, u_fix=None, uc_act=None, uc_fix=None, rho_c=0.0, gc_extra=None):
        """Backward vector pass, consensus solve, forward rollout.  rh (M,N,m), qh (M,N,n) are
        the per-iteration linear terms (cost + ADMM).  Returns X (M,N,n), U (M,N,m)."""
        M, N, n, m, Nc, s = self.M, self.N, self.n, self.m, self.Nc, self.s
        kff = np.zeros((M, N, m))
        p = np.zeros((M, n))
        z = np.zeros((M, m))
        for j in range(N - 1, Nc - 1, -1):
            S = self.Qs[:, j] + self.rho_x * np.eye(n) + self.Pst[:, j]
            B, A, W = self.B[:, j], self.A[:, j], self.Wst[:, j]
            eta = np.einsum("irt,it->ir", S, self.c[:, j]) + qh[:, j] + p
            h = rh[:, j] + np.einsum("itr,it->ir", B, eta) + np.einsum("itr,it->ir", W, self.c[:, j]) + z
            if u_fix is not None:
                uf = u_fix[:, j]
                h = h + np.einsum("irt,it->ir", self.Hfull[:, j], uf)
                eta = eta + np.einsum("irt,it->ir", self.SBW[:, j], uf)
            k = -np.einsum("irt,it->ir", self.Hinv[:, j], h)
            if u_fix is not None:
                k = k + u_fix[:, j]
            kff[:, j] = k
            p = np.einsum("itr,it->ir", A, eta) + np.einsum("itr,it->ir", self.K[:, j], h)
            z = -s[:, None] * k
        U = np.zeros((M, N, m))
        X = np.zeros((M, N, n))
        x = np.zeros((M, n))
        if Nc >= 1:
            D = Nc * m
            g0 = self._shared_grad(np.zeros((M, Nc, m)), qh, p, z).sum(0).reshape(-1)
            Hc = self.Hc_sum + rho_c * np.eye(D)
            if gc_extra is not None:
                g0 = g0 + gc_extra.reshape(-1)
            if uc_act is not None and uc_act.any():
                a = uc_act.reshape(-1) != 0
                beta = np.where(a, uc_fix.reshape(-1), 0.0)
                g0 = g0 + Hc @ beta
                Hm = np.where(a[:, None] | a[None, :], 0.0, Hc) + np.diag(a.astype(float))
                E = np.where(a[:, None] | a[None, :], 0.0, np.linalg.inv(Hm))
                uc = beta - E @ g0
            else:
                uc = -np.linalg.solve(Hc, g0)
            uc = uc.reshape(Nc, m)
            self._g0 = g0
            for j in range(Nc):
                U[:, j] = uc[j]
                x = np.einsum("irt,it->ir", self.A[:, j], x) + np.einsum("irt,it->ir", self.B[:, j], U[:, j]) + self.c[:, j]
                X[:, j] = x
        for j in range(Nc, N):
            u = np.einsum("irt,it->ir", self.K[:, j], x) + kff[:, j]
            if j >= 1:
                u = u + s[:, None] * np.einsum("irt,it->ir", self.Hinv[:, j], U[:, j - 1])
            x = np.einsum("irt,it->ir", self.A[:, j], x) + np.einsum("irt,it->ir", self.B[:, j], u) + self.c[:, j]
            U[:, j], X[:, j] = u, x
        return X, U

    # ------------------------------------------------------------------ KKT check (reduced gradient)
    def reduced_gradient(self, X, U, y_x=None):
        """Gradient of sum_i J_i w.r.t. every u_{i,j} with the dynamics eliminated (adjoint sweep
        with the TRUE cost, no ADMM terms).  State-box multipliers y_x enter the costate."""
        M, N, n, m, s = self.M, self.N, self.n, self.m, self.s
        g = np.zeros((M, N, m))
        a = np.zeros((M, n))
        for j in range(N - 1, -1, -1):
            a_new = np.einsum("irt,it->ir", self.Qs[:, j], X[:, j]) + self.q[:, j]
            if y_x is not None:
                a_new = a_new + y_x[:, j]
            if j < N - 1:
                a_new = a_new + np.einsum("itr,it->ir", self.A[:, j + 1], a)
            a = a_new
            gu = np.einsum("irt,it->ir", self.Rs[:, j], U[:, j]) + self.r[:, j] + np.einsum("itr,it->ir", self.B[:, j], a)
            if j >= 1:
                gu = gu + s[:, None] * (U[:, j] - U[:, j - 1])
            if j < N - 1:
                gu = gu + s[:, None] * (U[:, j] - U[:, j + 1])
            g[:, j] = gu
        return g

    def kkt_violation(self, X, U, y_x=None):
        """max KKT violation of (X,U) for the box QP, relative; also returns duals of control box."""
        Nc = self.Nc
        g = self.reduced_gradient(X, U, y_x)
        if Nc >= 1:
            g[:, :Nc] = g[:, :Nc].sum(0, keepdims=True)  # shared block: summed gradient
        scale = max(1.0, np.abs(self.r).max(), np.abs(self.q).max())
        zs = max(1.0, np.abs(U).max(), np.abs(X).max())
        at_u = U >= self.uu - 1e-12 * zs
        at_l = U <= self.lu + 1e-12 * zs
        # stationarity: g = 0 strictly inside, g <= 0 at upper, g >= 0 at lower
        viol = np.where(at_u & at_l, 0.0, np.where(at_u, np.maximum(g, 0), np.where(at_l, np.maximum(-g, 0), np.abs(g))))
        stat = viol.max() / scale
        pu = max(np.maximum(U - self.uu, 0).max(), np.maximum(self.lu - U, 0).max()) / zs
        px = max(np.maximum(X - self.ux, 0).max(), np.maximum(self.lx - X, 0).max()) / zs
        return max(stat, pu, px), -g

    # ------------------------------------------------------------------ polish
    def polish(self, act_u, st: Settings):
        """Primal-dual active-set refinement on the control box (state box must be inactive)."""
        M, N, m, Nc = self.M, self.N, self.m, self.Nc
        zero_x = np.zeros((M, N, self.n))
        for it in range(st.polish_iters):
            self.stats["polish_attempts"] += 1
            act_u = act_u.copy()
            act_u[:, :Nc] = act_u[0:1, :Nc]
            beta = np.where(act_u > 0, self.uu, np.where(act_u < 0, self.lu, 0.0))
            beta = np.where(act_u != 0, beta, 0.0)
            self.factor(0.0, 0.0, act_u=act_u)
            uc_act = act_u[0, :Nc] if Nc >= 1 else None
            uc_fix = beta[0, :Nc] if Nc >= 1 else None
            X, U = self.solve_lq(self.r, self.q + zero_x, u_fix=beta, uc_act=uc_act, uc_fix=uc_fix)
            viol, mu = self.kkt_violation(X, U)
            if viol <= st.kkt_tol:
                return True, X, U, mu, act_u
            if self.has_xbox:
                zs = max(1.0, np.abs(X).max())
                if max(np.maximum(X - self.ux, 0).max(), np.maximum(self.lx - X, 0).max()) > st.kkt_tol * zs:
                    return False, X, U, mu, act_u  # a state bound wants to be active: ADMM's job
            # PDAS update: mu = -g is the multiplier estimate (mu>0 pushes against upper bound)
            new = np.zeros_like(act_u)
            new[(mu + (U - self.uu)) > 0] = 1
            new[(mu + (U - self.lu)) < 0] = -1
            new[(new > 0) & ~np.isfinite(self.uu)] = 0
            new[(new < 0) & ~np.isfinite(self.lu)] = 0
            if np.array_equal(new, act_u):
                return False, X, U, mu, act_u
            act_u = new
        return False, X, U, mu, act_u

    # ------------------------------------------------------------------ ADMM driver
    def solve(self, st: Settings = None, state: State = None):
        st = st or Settings()
        M, N, n, m, Nc = self.M, self.N, self.n, self.m, self.Nc
        if state is None or state.w_u is None:
            state = State(np.zeros((M, N, m)), np.zeros((M, N, m)), np.zeros((M, N, n)), np.zeros((M, N, n)), st.rho)
        w_u, y_u, w_x, y_x, rho = state.w_u, state.y_u, state.w_x, state.y_x, state.rho
        constrained = self.has_ubox or self.has_xbox
        if not constrained:
            self.factor(0.0, 0.0)
            X, U = self.solve_lq(self.r, self.q)
            self.stats["polish_ok"] = True
            return X, U, state
        act_guess = np.zeros((M, N, m), dtype=np.int8)
        act_guess[(self.uu - w_u) < y_u] = 1
        act_guess[(w_u - self.lu) < -y_u] = -1
        X = U = None
        it = 0
        need_factor = True
        while True:
            # ---- polish attempt (skipped while a state bound is active in the ADMM iterate)
            x_active = self.has_xbox and np.any(y_x != 0.0)
            if st.polish and not x_active:
                ok, Xp, Up, mu, act = self.polish(act_guess, st)
                need_factor = True
                if ok:
                    self.stats["polish_ok"] = True
                    # hand consistent duals back for the next warm start
                    state = State(np.clip(Up, self.lu, self.uu), np.where(act != 0, mu, 0.0), np.clip(Xp, self.lx, self.ux),
                                  np.zeros_like(y_x), rho)
                    self.stats["admm_iters"] = it
                    return Xp, Up, state
            if it >= st.max_iter:
                break
            # ---- a block of ADMM iterations
            rho_u = rho if self.has_ubox else 0.0
            rho_x = rho if self.has_xbox else 0.0
            if need_factor:
                self.factor(rho_u, rho_x)
                need_factor = False
            for _ in range(st.check_every):
                rh = self.r + rho_u * (-w_u) + (y_u if self.has_ubox else 0.0)
                rh[:, :Nc] = self.r[:, :Nc]  # shared block penalty enters the Schur system once
                qh = self.q + rho_x * (-w_x) + (y_x if self.has_xbox else 0.0)
                gc_extra = (-rho_u * w_u[0, :Nc] + y_u[0, :Nc]) if (Nc >= 1 and self.has_ubox) else None
                X, U = self.solve_lq(rh, qh, rho_c=rho_u if Nc >= 1 else 0.0, gc_extra=gc_extra)
                it += 1
                rp = rd = 0.0
                if self.has_ubox:
                    zr = st.alpha * U + (1 - st.alpha) * w_u
                    wn = np.clip(zr + y_u / rho, self.lu, self.uu)
                    y_u = y_u + rho * (zr - wn)
                    rp, rd = max(rp, np.abs(U - wn).max()), max(rd, rho * np.abs(wn - w_u).max())
                    w_u = wn
                if self.has_xbox:
                    zr = st.alpha * X + (1 - st.alpha) * w_x
                    wn = np.clip(zr + y_x / rho, self.lx, self.ux)
                    y_x = y_x + rho * (zr - wn)
                    rp, rd = max(rp, np.abs(X - wn).max()), max(rd, rho * np.abs(wn - w_x).max())
                    w_x = wn
            sp = max(np.abs(U).max(), np.abs(w_u).max(), np.abs(X).max() if self.has_xbox else 0.0, 1e-12)
            sd = max(np.abs(y_u).max(), np.abs(y_x).max(), np.abs(self.r).max(), np.abs(self.q).max(), 1e-12)
            self.stats["r_prim"], self.stats["r_dual"] = rp, rd
            if rp <= st.eps_abs + st.eps_rel * sp and rd <= st.eps_abs + st.eps_rel * sd:
                break
            if st.adapt_rho:
                est = rho * np.sqrt((rp / sp) / max(rd / sd, 1e-30))
                est = float(np.clip(est, 1e-6, 1e6))
                if est > rho * st.adapt_tol or est < rho / st.adapt_tol:
                    rho, need_factor = est, True
            act_guess = np.zeros((M, N, m), dtype=np.int8)
            act_guess[(self.uu - w_u) < y_u] = 1
            act_guess[(w_u - self.lu) < -y_u] = -1
        self.stats["admm_iters"] = it
        state = State(w_u, y_u, w_x, y_x, rho)
        return X, U, state


def lqp_solve(*args, settings: Settings = None, state: State = None, **kw):
    sol = RiccatiADMM(*args, **kw)
    X, U, state = sol.solve(settings, state)
    return X, U, dict(stats=sol.stats, state=state, solver=sol)

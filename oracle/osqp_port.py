"""Restatement of the OSQP 0.6.2 ADMM.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference solves its QP with **OSQP.jl 0.8.0 / OSQP_jll 0.600.200 (OSQP C 0.6.2)**
(`PMPC.jl/Manifest.toml:294-304`; call sites `PMPC.jl/src/osqp_solver.jl:40,48,50,67`),
a third-party dependency whose source is NOT under `/root/reference`.  This file
restates its *published* algorithm (Stellato et al., "OSQP: an operator splitting
solver for quadratic programs", Math. Prog. Comp. 2020, Alg. 1, Sec. 5) with the
0.6.2 defaults the reference runs with (it passes no settings through the C ABI,
`c_interface.jl:50-71`):

    rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=eps_rel=1e-3, max_iter=4000,
    scaling=10 (Ruiz), adaptive_rho (tolerance 5), check_termination=25, polish off,
    equality rows get 1e3*rho.

The stacked constraint is `Aa=[A;G]`, `la=[b;l]`, `ua=[b;u]` exactly as
`osqp_solver.jl:20-32` builds it.  The linear system is factored with SciPy's
SuperLU instead of QDLDL+AMD, so timings are "restated CPU baseline, not PMPC.jl".
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_EQ_FACTOR = 1e3
MIN_SCALING, MAX_SCALING = 1e-4, 1e4


@dataclass
class OSQPResult:
    x: np.ndarray
    y: np.ndarray
    iters: int
    status: str
    rho_updates: int
    r_prim: float
    r_dual: float
    setup_time: float
    solve_time: float


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _ruiz(P, q, A, iters):
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, A, q = P.copy().tocsc(), A.copy().tocsc(), q.copy()
    for _ in range(iters):
        colP = np.abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
        colA = np.abs(A).max(axis=0).toarray().ravel() if A.nnz else np.zeros(n)
        rowA = np.abs(A).max(axis=1).toarray().ravel() if A.nnz else np.zeros(m)
        d = 1.0 / np.sqrt(_limit(np.maximum(colP, colA)))
        e = 1.0 / np.sqrt(_limit(rowA))
        Dm, Em = sp.diags(d), sp.diags(e)
        P = (Dm @ P @ Dm).tocsc()
        A = (Em @ A @ Dm).tocsc()
        q = d * q
        D *= d
        E *= e
        colP = np.abs(P).max(axis=0).toarray().ravel() if P.nnz else np.zeros(n)
        cost = max(colP.mean() if n else 0.0, np.abs(q).max(initial=0.0))
        cost = 1.0 if cost < MIN_SCALING else min(cost, MAX_SCALING)
        ct = 1.0 / cost
        P = P * ct
        q = q * ct
        c *= ct
    return P.tocsc(), q, A.tocsc(), D, E, c


def osqp_solve(
    prob,
    rho=0.1,
    sigma=1e-6,
    alpha=1.6,
    eps_abs=1e-3,
    eps_rel=1e-3,
    max_iter=4000,
    scaling=10,
    adaptive_rho=True,
    adaptive_rho_tolerance=5.0,
    adaptive_rho_fraction=0.4,
    check_termination=25,
) -> OSQPResult:
    """Solve ``prob`` (an ``oracle.lqp.LQP``) the way `osqp_solver.jl:34-72` does."""
    t0 = time.perf_counter()
    P0, q0 = prob.P, prob.q
    Aa0 = sp.vstack([prob.A, prob.G], format="csc")  # osqp_solver.jl:20-32
    la0 = np.concatenate([prob.b, prob.l])
    ua0 = np.concatenate([prob.b, prob.u])
    n, m = P0.shape[0], Aa0.shape[0]
    if scaling:
        P, q, A, D, E, c = _ruiz(P0, q0, Aa0, scaling)
    else:
        P, q, A, D, E, c = P0, q0.copy(), Aa0, np.ones(n), np.ones(m), 1.0
    l, u = E * la0, E * ua0
    Dinv, Einv = 1.0 / D, 1.0 / E
    is_eq = (ua0 - la0) < 1e-4  # OSQP: RHO_TOL
    is_free = (la0 < -1e26) & (ua0 > 1e26)  # OSQP_INFTY * MIN_SCALING

    def rho_vec_of(r):
        v = np.full(m, r)
        v[is_eq] = RHO_EQ_FACTOR * r
        v[is_free] = RHO_MIN
        return v

    AT = A.T.tocsc()
    Psig = (P + sigma * sp.identity(n)).tocsc()

    def factor(rv):
        K = sp.bmat([[Psig, AT], [A, -sp.diags(1.0 / rv)]], format="csc")
        return spla.splu(K)

    rv = rho_vec_of(rho)
    lu = factor(rv)
    setup_time = time.perf_counter() - t0

    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    t1 = time.perf_counter()
    status, rho_updates, adapt_interval = "max_iter_reached", 0, 0
    rp = rd = np.inf
    it = 0
    for it in range(1, max_iter + 1):
        sol = lu.solve(np.concatenate([sigma * x - q, z - y / rv]))
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, l, u)
        y = y + rv * (zr - zn)
        z = zn
        # automatic adaptive-rho interval (OSQP: a multiple of check_termination once the time
        # spent iterating reaches adaptive_rho_fraction * setup_time)
        if adaptive_rho and adapt_interval == 0:
            if time.perf_counter() - t1 > adaptive_rho_fraction * setup_time:
                k = check_termination if check_termination else 25
                adapt_interval = max(k, int(round(it / k)) * k)
        check = check_termination and it % check_termination == 0
        adapt = adaptive_rho and adapt_interval and it % adapt_interval == 0
        if check or adapt or it == max_iter:
            Ax, Px, Aty = A @ x, P @ x, AT @ y
            rp = np.abs(Einv * (Ax - z)).max(initial=0.0)
            rd = np.abs(Dinv * (Px + q + Aty)).max(initial=0.0) / c
            n_prim = max(np.abs(Einv * Ax).max(initial=0.0), np.abs(Einv * z).max(initial=0.0))
            n_dual = max(
                np.abs(Dinv * Px).max(initial=0.0),
                np.abs(Dinv * Aty).max(initial=0.0),
                np.abs(Dinv * q).max(initial=0.0),
            ) / c
            if check and rp <= eps_abs + eps_rel * n_prim and rd <= eps_abs + eps_rel * n_dual:
                status = "solved"
                break
            if adapt:
                # scaled-space residual norms as in OSQP's compute_rho_estimate
                rps = np.abs(Ax - z).max(initial=0.0)
                rds = np.abs(Px + q + Aty).max(initial=0.0)
                nps = max(np.abs(Ax).max(initial=0.0), np.abs(z).max(initial=0.0))
                nds = max(np.abs(Px).max(initial=0.0), np.abs(Aty).max(initial=0.0), np.abs(q).max(initial=0.0))
                est = rho * np.sqrt((rps / (nps + 1e-10)) / (rds / (nds + 1e-10) + 1e-30))
                est = float(np.clip(est, RHO_MIN, RHO_MAX))
                if est > rho * adaptive_rho_tolerance or est < rho / adaptive_rho_tolerance:
                    rho = est
                    rv = rho_vec_of(rho)
                    lu = factor(rv)
                    rho_updates += 1
    solve_time = time.perf_counter() - t1
    return OSQPResult(D * x, (E * y) / c, it, status, rho_updates, float(rp), float(rd), setup_time, solve_time)


def lqp_solve_osqp(*args, osqp_settings=None, **kw):
    """`main.jl:115-171` with the reference's solver restated: returns X, U, result."""
    from .lqp import build_lqp, split_vars

    prob = build_lqp(*args, **kw)
    res = osqp_solve(prob, **(osqp_settings or {}))
    X, U = split_vars(prob, res.x)
    return X, U, res

"""Oracle for the consensus QP of one SCP iteration.  TEST INFRASTRUCTURE ONLY.

Restates (vectorised NumPy, not a transcription) what the reference builds in

* ``PMPC.jl/src/lqp_utils.jl:2-216``   ``lqp_repr_Pq``   (Hessian / linear term)
* ``PMPC.jl/src/lqp_utils.jl:219-303`` ``lqp_repr_Ab``   (linearised dynamics rows)
* ``PMPC.jl/src/lqp_utils.jl:306-393`` ``lqp_repr_Gla``  (box rows)
* ``PMPC.jl/src/lqp_utils.jl:395-423`` ``split_lqp_vars``
* ``PMPC.jl/src/c_interface.jl:56-70`` NaN sentinels, ``main.jl:127-128`` ``Nc<0 -> N``

and solves ``min 1/2 z'Pz + q'z  s.t.  Az=b, l<=Gz<=u`` *exactly* (primal-dual
active set on the sparse KKT system, SciPy SuperLU), returning a
solver-independent KKT certificate.  The reference hands the same matrices to
OSQP (``osqp_solver.jl:34-72``); since the QP is strictly convex its minimiser is
unique, so any exact solver defines the same answer.

All arrays use the Python layout of ``pmpc.solve``: vectors ``(M, N, d)``,
matrices ``(M, N, rows, cols)`` (``static_backend.py:223-241`` only re-views them).

Variable order (``README.md:232-239`` of the reference):
``z = [u_c (Nc*m) ; u_free particle-major (M*Nf*m) ; x particle-major (M*N*n)]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------- data
@dataclass
class LQP:
    """The assembled QP plus index maps."""

    P: sp.csc_matrix  # symmetric (upper triangle of the reference's P mirrored, as OSQP reads it)
    q: np.ndarray
    A: sp.csc_matrix
    b: np.ndarray
    G: sp.csc_matrix
    l: np.ndarray
    u: np.ndarray
    M: int
    N: int
    n: int
    m: int
    Nc: int
    iu: np.ndarray  # (M, N, m) index of u[i, j, :] in z
    ix: np.ndarray  # (M, N, n) index of x[i, j, :] in z

    @property
    def nz(self):
        return self.q.size


def _nan_absent(*arrs):
    """c_interface.jl:56-70 — any NaN in a family disables the whole family."""
    if any(a is None for a in arrs):
        return True
    return any(np.isnan(np.asarray(a)).any() for a in arrs)


def index_maps(M, N, n, m, Nc):
    Nf = N - Nc
    iu = np.empty((M, N, m), dtype=np.int64)
    t = np.arange(m)
    for j in range(Nc):
        iu[:, j, :] = m * j + t
    i = np.arange(M)[:, None, None]
    jf = np.arange(Nf)[None, :, None]
    iu[:, Nc:, :] = Nc * m + Nf * m * i + m * jf + t
    off = Nc * m + M * Nf * m
    ix = off + N * n * i + n * np.arange(N)[None, :, None] + np.arange(n)
    return iu, ix


def build_lqp(
    x0,
    f,
    fx,
    fu,
    X_prev,
    U_prev,
    Q,
    R,
    X_ref,
    U_ref,
    lx=None,
    ux=None,
    lu=None,
    uu=None,
    reg_x=0.0,
    reg_u=0.0,
    slew_reg=None,
    slew_reg0=None,
    slew_um1=None,
    Nc=-1,
    lin_rows=None,
) -> LQP:
    f = np.asarray(f, dtype=np.float64)
    M, N, n = f.shape
    m = np.asarray(fu).shape[-1]
    fx, fu = np.asarray(fx, float).reshape(M, N, n, n), np.asarray(fu, float).reshape(M, N, n, m)
    Q, R = np.asarray(Q, float).reshape(M, N, n, n), np.asarray(R, float).reshape(M, N, m, m)
    X_prev, X_ref = np.asarray(X_prev, float).reshape(M, N, n), np.asarray(X_ref, float).reshape(M, N, n)
    U_prev, U_ref = np.asarray(U_prev, float).reshape(M, N, m), np.asarray(U_ref, float).reshape(M, N, m)
    Nc_in = int(Nc)
    Nc = N if Nc_in < 0 else Nc_in  # main.jl:127-128
    assert 0 <= Nc <= N
    reg_x = np.broadcast_to(np.asarray(reg_x, float), (M,))  # main.jl:19-21
    reg_u = np.broadcast_to(np.asarray(reg_u, float), (M,))
    # sentinels (c_interface.jl:64-70, main.jl:23-27)
    s = np.zeros(M) if _nan_absent(slew_reg) else np.broadcast_to(np.asarray(slew_reg, float), (M,))
    if _nan_absent(slew_reg0, slew_um1):
        s0, um1 = np.zeros(M), np.zeros((M, m))
    else:
        s0 = np.broadcast_to(np.asarray(slew_reg0, float), (M,))
        um1 = np.asarray(slew_um1, float).reshape(M, m)
    has_ubox = not _nan_absent(lu, uu)
    has_xbox = not _nan_absent(lx, ux)

    iu, ix = index_maps(M, N, n, m, Nc)
    nz = Nc * m + M * ((N - Nc) * m + N * n)

    rows, cols, vals = [], [], []

    def add(r, c, v):
        rows.append(np.broadcast_to(r, v.shape).ravel())
        cols.append(np.broadcast_to(c, v.shape).ravel())
        vals.append(np.asarray(v, float).ravel())

    # ---- P: control blocks (lqp_utils.jl:17-103).  Duplicates are summed, which is exactly the
    # un-weighted sum over particles on the consensus block (lqp_utils.jl:27-40).
    add(iu[:, :, :, None], iu[:, :, None, :], R)
    dg = np.empty((M, N))
    dg[:] = reg_u[:, None] + 2.0 * s[:, None]
    dg[:, N - 1] = reg_u + s  # j == N
    dg[:, 0] = reg_u + s0 + s  # j == 1 wins over j == N when N == 1 (lqp_utils.jl:31-37)
    add(iu, iu, np.broadcast_to(dg[:, :, None], (M, N, m)))
    if N > 1:
        off = np.broadcast_to(-s[:, None, None], (M, N - 1, m))
        add(iu[:, :-1, :], iu[:, 1:, :], off)
        add(iu[:, 1:, :], iu[:, :-1, :], off)
    # ---- P: state blocks (lqp_utils.jl:106-160)
    add(ix[:, :, :, None], ix[:, :, None, :], Q)
    add(ix, ix, np.broadcast_to(reg_x[:, None, None], (M, N, n)))
    P_full = sp.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nz, nz)
    ).tocsc()
    # OSQP reads the upper triangle only
    Pu = sp.triu(P_full, format="csc")
    P = (Pu + sp.triu(P_full, k=1, format="csc").T).tocsc()

    # ---- q (lqp_utils.jl:164-214)
    q = np.zeros(nz)
    qu = -reg_u[:, None, None] * U_prev - np.einsum("ijrt,ijt->ijr", R, U_ref)
    np.add.at(q, iu.ravel(), qu.ravel())
    if Nc >= 1:
        # `q[1:udim] .+= sum_i -slew_reg0_i * slew_um1_i`; with Nc == 0 the free-control loop
        # overwrites these entries (lqp_utils.jl:165 vs :187) -- reproduced, see SURVEY App. B.6
        q[:m] += -(s0[:, None] * um1).sum(0)
    qx = -reg_x[:, None, None] * X_prev - np.einsum("ijrt,ijt->ijr", Q, X_ref)
    q[ix.ravel()] = qx.ravel()

    # ---- A, b (lqp_utils.jl:219-303)
    rows, cols, vals = [], [], []
    er = (N * n * np.arange(M)[:, None, None] + n * np.arange(N)[None, :, None] + np.arange(n))  # (M,N,n)
    add(er[:, :, :, None], iu[:, :, None, :], fu)
    add(er, ix, -np.ones((M, N, n)))
    if N > 1:
        add(er[:, 1:, :, None], ix[:, :-1, None, :], fx[:, 1:])
    A = sp.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(M * N * n, nz)
    ).tocsc()
    b = -f + np.einsum("ijrt,ijt->ijr", fu, U_prev)
    if N > 1:
        b[:, 1:] += np.einsum("ijrt,ijt->ijr", fx[:, 1:], X_prev[:, :-1])
    b = b.reshape(-1)

    # ---- G, l, u (lqp_utils.jl:306-393); shared block bounds come from particle 1 (:329-330)
    gcols, gl, gu = [], [], []
    if has_ubox:
        lu, uu = np.asarray(lu, float).reshape(M, N, m), np.asarray(uu, float).reshape(M, N, m)
        gcols.append(iu[0, :Nc].ravel())
        gl.append(lu[0, :Nc].ravel())
        gu.append(uu[0, :Nc].ravel())
        gcols.append(iu[:, Nc:].ravel())
        gl.append(lu[:, Nc:].ravel())
        gu.append(uu[:, Nc:].ravel())
    if has_xbox:
        lx, ux = np.asarray(lx, float).reshape(M, N, n), np.asarray(ux, float).reshape(M, N, n)
        gcols.append(ix.ravel())
        gl.append(lx.ravel())
        gu.append(ux.ravel())
    if gcols:
        gc = np.concatenate(gcols)
        G = sp.coo_matrix((np.ones(gc.size), (np.arange(gc.size), gc)), shape=(gc.size, nz)).tocsc()
        l, u = np.concatenate(gl), np.concatenate(gu)
    else:
        G = sp.csc_matrix((0, nz))
        l, u = np.zeros(0), np.zeros(0)
    if lin_rows is not None:
        # general stage-local linear rows  Gu[i,j,k,:] u_{i,j} + Gx[i,j,k,:] x_{i,j} <= h[i,j,k]  (storage stage j: the control
        # and the state it produces share the index, README.md:232-239) -- what the reference admits as the `l` rows of an
        # `extra_cstrs` tuple (main.jl:293-316, cone_utils.jl:97-150: G_left z <= h), appended below the box rows
        Gu, Gx, hh = (np.asarray(a, float) for a in lin_rows)
        L = hh.shape[-1]
        Gu, Gx, hh = Gu.reshape(M, N, L, m), Gx.reshape(M, N, L, n), hh.reshape(M, N, L)
        rr = np.arange(M * N * L).reshape(M, N, L)
        Gl = sp.coo_matrix(
            (np.concatenate([Gu.ravel(), Gx.ravel()]),
             (np.concatenate([np.broadcast_to(rr[..., None], Gu.shape).ravel(), np.broadcast_to(rr[..., None], Gx.shape).ravel()]),
              np.concatenate([np.broadcast_to(iu[:, :, None, :], Gu.shape).ravel(), np.broadcast_to(ix[:, :, None, :], Gx.shape).ravel()]))),
            shape=(M * N * L, nz)).tocsc()
        keep = np.isfinite(hh.ravel())
        Gl = Gl[np.flatnonzero(keep)]
        G = sp.vstack([G, Gl], format="csc")
        l = np.concatenate([l, np.full(int(keep.sum()), -np.inf)])
        u = np.concatenate([u, hh.ravel()[keep]])
    return LQP(P, q, A, b, G, l, u, M, N, n, m, Nc, iu, ix)


def split_vars(prob: LQP, z):
    """lqp_utils.jl:395-423 — consensus block broadcast to every particle."""
    return z[prob.ix], z[prob.iu]


# --------------------------------------------------------------------------- exact solve
@dataclass
class Certificate:
    stationarity: float
    primal_eq: float
    primal_box: float
    dual_sign: float
    complementarity: float
    objective: float
    iterations: int

    @property
    def worst(self):
        return max(self.stationarity, self.primal_eq, self.primal_box, self.dual_sign, self.complementarity)


def kkt_certificate(prob: LQP, z, nu, mu, iterations=0) -> Certificate:
    """Relative KKT residuals of (z, nu, mu) for the QP (SURVEY.md 8c 'Consequence')."""
    P, q, A, b, G, l, u = prob.P, prob.q, prob.A, prob.b, prob.G, prob.l, prob.u
    Pz = P @ z
    g = Pz + q + A.T @ nu + (G.T @ mu if mu.size else 0.0)
    sc = max(1.0, np.abs(q).max(initial=0.0), np.abs(Pz).max(initial=0.0))
    stat = np.abs(g).max(initial=0.0) / sc
    peq = np.abs(A @ z - b).max(initial=0.0) / max(1.0, np.abs(b).max(initial=0.0))
    if mu.size:
        Gz = G @ z
        zs = max(1.0, np.abs(z).max())
        pbox = max(np.maximum(Gz - u, 0).max(), np.maximum(l - Gz, 0).max()) / zs
        at_u, at_l = np.abs(Gz - u) <= 1e-9 * zs, np.abs(Gz - l) <= 1e-9 * zs
        ms = max(1.0, np.abs(mu).max())
        # mu > 0 only at the upper bound, mu < 0 only at the lower bound
        bad = np.where(mu > 0, ~at_u, np.where(mu < 0, ~at_l, False))
        dsign = np.abs(mu[bad]).max(initial=0.0) / ms
        gap = np.minimum(np.abs(Gz - u), np.abs(Gz - l))
        comp = np.abs(np.where(mu != 0, gap, 0.0) * mu).max(initial=0.0) / (ms * zs)
    else:
        pbox = dsign = comp = 0.0
    obj = 0.5 * z @ Pz + q @ z
    return Certificate(float(stat), float(peq), float(pbox), float(dsign), float(comp), float(obj), iterations)


def solve_exact(prob: LQP, max_iter: int = 200, act0=None):
    """Primal-dual active-set (semismooth Newton) solve on the sparse KKT system.

    Returns ``(z, nu, mu, cert)``.  Globalised by nothing: for the strictly convex,
    identity-row box QPs of this path it terminates in a handful of iterations; the
    certificate is what makes the result trustworthy, not the method.
    """
    P, q, A, b, G, l, u = prob.P, prob.q, prob.A, prob.b, prob.G, prob.l, prob.u
    nz, ne, ni = prob.nz, A.shape[0], G.shape[0]
    Gr = G.tocsr()  # box rows are unit selection rows; `lin_rows` add general ones
    act = np.zeros(ni, dtype=np.int8) if act0 is None else act0.copy()
    z = np.zeros(nz)
    nu = np.zeros(ne)
    mu = np.zeros(ni)
    seen = set()
    for it in range(1, max_iter + 1):
        idx = np.flatnonzero(act)
        na = idx.size
        Ga = Gr[idx].tocsc() if na else sp.csc_matrix((0, nz))
        rhs_a = np.where(act[idx] > 0, u[idx], l[idx])
        # tiny regularisation on the multiplier blocks + iterative refinement (OSQP-polish style)
        delta = 1e-10
        K = sp.bmat(
            [
                [P, A.T, Ga.T],
                [A, -delta * sp.identity(ne), None],
                [Ga, None, -delta * sp.identity(na)],
            ],
            format="csc",
        )
        K0 = sp.bmat([[P, A.T, Ga.T], [A, None, None], [Ga, None, None]], format="csc")
        rhs = np.concatenate([-q, b, rhs_a])
        lu_ = spla.splu(K)
        sol = lu_.solve(rhs)
        for _ in range(5):
            sol += lu_.solve(rhs - K0 @ sol)
        z, nu = sol[:nz], sol[nz : nz + ne]
        mu = np.zeros(ni)
        mu[idx] = sol[nz + ne :]
        if ni == 0:
            break
        Gz = Gr @ z
        c = 1.0
        new = np.zeros(ni, dtype=np.int8)
        new[(mu + c * (Gz - u)) > 0] = 1
        new[(mu + c * (Gz - l)) < 0] = -1
        # infinite bounds can never be active
        new[(new > 0) & ~np.isfinite(u)] = 0
        new[(new < 0) & ~np.isfinite(l)] = 0
        if np.array_equal(new, act):
            break
        key = new.tobytes()
        if key in seen:  # cycling: leave it to the certificate / fallback
            act = new
            break
        seen.add(key)
        act = new
    cert = kkt_certificate(prob, z, nu, mu, iterations=it)
    return z, nu, mu, cert


def lqp_solve(*args, **kw):
    """Oracle counterpart of ``PMPC.jl/src/main.jl:115-171`` ``lqp_solve`` (no coerce / weights:
    neither is reachable through the C ABI, c_interface.jl:50-71).

    Returns ``X (M,N,n)``, ``U (M,N,m)``, ``info`` (certificate, objective, multipliers)."""
    require = kw.pop("require_certificate", 1e-8)
    prob = build_lqp(*args, **kw)
    z, nu, mu, cert = solve_exact(prob)
    if cert.worst > require:
        # fall back on the restated OSQP algorithm at a tight tolerance, then re-polish
        from .osqp_port import osqp_solve

        res = osqp_solve(prob, eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)
        act = np.zeros(prob.G.shape[0], dtype=np.int8)
        yb = res.y[prob.A.shape[0] :]
        act[yb > 1e-7] = 1
        act[yb < -1e-7] = -1
        z, nu, mu, cert = solve_exact(prob, act0=act)
        if cert.worst > require:
            raise RuntimeError(f"oracle could not certify the QP solution: {cert}")
    X, U = split_vars(prob, z)
    return X, U, dict(cert=cert, z=z, nu=nu, mu=mu, prob=prob)

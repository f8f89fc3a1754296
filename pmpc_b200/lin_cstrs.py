"""General stage-local linear constraints  ``g' [u_j; x_j] <= h``  (north_star: "box and linear control/state
constraints"; BASELINE config 4's ``n_lin`` rows).

The reference admits them only as the ``l`` rows of an ``extra_cstrs`` tuple ``(l, q, e, G_left, G_right, h, c_left,
c_right)`` handed to its cone route (``pmpc/scp_mpc.py:353-361``, ``PMPC.jl/src/main.jl:293-316``,
``cone_utils.jl:97-150``: ``G_left z <= h`` on the variable vector ``z = [controls; states]``, ``README.md:219-239``).
Here they are served by the QP route with NO new kernel, by a change of variables: the row values become extra states

    s_j = Gu_j u_j + Gx_j x_j ,        x_j = f_j + fx_j (x_{j-1} - xbar) + fu_j (u_j - ubar)
        = (Gu_j ubar_j + Gx_j f_j)  +  (Gx_j fx_j) (x_{j-1} - xbar)  +  (Gu_j + Gx_j fu_j) (u_j - ubar_j)

i.e. ``L`` more rows in ``f, fx, fu`` (zero columns for the new states), zero cost on them (their ``Q`` diagonal is
``-reg_x``, which cancels the proximal weight the back end adds to every state, ``lqp_utils.jl:130-142``; their
``X_prev`` is kept at zero so the proximal term has no linear part either), and the rows turn into the STATE BOX
``s_j <= h_j`` that the device solver already handles (Riccati-ADMM with fused projection / dual update).  The
minimiser over (x, u) is unchanged: the new states are determined by (x, u) and cost nothing.

Index convention (the reference's variable layout): row k of stage j couples ``U[i, j, :]`` with ``X[i, j + 1, :]`` of the
returned trajectory — the control and the state it produces share the storage index.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def from_extra_cstrs(extra_cstrs: Sequence, M: int, N: int, n: int, m: int, Nc: int):
    """Reference-format constraint tuples -> ``(Gu (M,N,L,m), Gx (M,N,L,n), h (M,N,L))`` (rows padded with h = +inf).

    Only what the QP route can represent is accepted: linear rows (``l > 0``, no second-order / exponential cones, no new
    variables, no cost augmentation) whose non-zeros stay within ONE (particle, stage) block of the variable vector.
    Anything else raises ``NotImplementedError`` naming the offending piece."""
    Nc = N if Nc < 0 else int(Nc)
    if M == 1:
        Nc = 0  # one particle: [u_1 .. u_N ; x_1 .. x_N] whatever Nc says (README.md:232-235)
    Nf = N - Nc
    nu_tot = Nc * m + M * Nf * m
    nz = nu_tot + M * N * n
    per = [[[] for _ in range(N)] for _ in range(M)]
    for tup in extra_cstrs:
        l, q, e, G_left, G_right, h, c_left, c_right = tup
        if len(list(q)) > 0 or int(e) != 0:
            raise NotImplementedError("extra_cstrs: second-order / exponential cone rows need the reference's cone route, "
                                      "which pmpc_b200 does not provide (DESIGN.md, out of scope)")
        G_right = np.asarray(G_right, dtype=np.float64)
        if G_right.size and G_right.reshape(int(l), -1).shape[1] > 0:
            raise NotImplementedError("extra_cstrs: rows that introduce new variables (G_right) are not supported")
        if np.any(np.asarray(c_left, dtype=np.float64) != 0.0) or np.size(c_right) > 0:
            raise NotImplementedError("extra_cstrs: cost augmentation (c_left / c_right) is not supported; use lin_cost_fn")
        G = np.asarray(G_left.todense() if hasattr(G_left, "todense") else G_left, dtype=np.float64).reshape(int(l), -1)
        h = np.asarray(h, dtype=np.float64).reshape(-1)
        if G.shape[1] > nz:
            raise ValueError(f"extra_cstrs: G_left has {G.shape[1]} columns, the variable vector has {nz}")
        for k in range(int(l)):
            cols = np.flatnonzero(G[k])
            if cols.size == 0:
                continue
            owner = set()
            gu, gx = np.zeros(m), np.zeros(n)
            for c in cols:
                if c < Nc * m:
                    if M > 1:
                        raise NotImplementedError("extra_cstrs: rows on the shared (consensus) controls of a multi-particle problem")
                    i, j, a = 0, c // m, c % m
                    gu[a] = G[k, c]
                elif c < nu_tot:
                    cc = c - Nc * m
                    i, jj, a = cc // (Nf * m), (cc % (Nf * m)) // m, cc % m
                    j = Nc + jj
                    gu[a] = G[k, c]
                else:
                    cc = c - nu_tot
                    i, j, r = cc // (N * n), (cc % (N * n)) // n, cc % n
                    gx[r] = G[k, c]
                owner.add((i, j))
            if len(owner) != 1:
                raise NotImplementedError("extra_cstrs: a row couples several stages / particles; only stage-local rows "
                                          "(one control u_j and the state it produces) are supported")
            (i, j), = owner
            per[i][j].append((gu, gx, h[k]))
    L = max(1, max(len(per[i][j]) for i in range(M) for j in range(N)))
    Gu, Gx, hh = np.zeros((M, N, L, m)), np.zeros((M, N, L, n)), np.full((M, N, L), np.inf)
    for i in range(M):
        for j in range(N):
            for k, (gu, gx, hv) in enumerate(per[i][j]):
                Gu[i, j, k], Gx[i, j, k], hh[i, j, k] = gu, gx, hv
    return Gu, Gx, hh


def normalise(lin, M: int, N: int, n: int, m: int) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """``(Gu, Gx, h)`` broadcast to ``(M,N,L,m), (M,N,L,n), (M,N,L)``; ``None`` stays ``None``."""
    if lin is None:
        return None
    Gu, Gx, h = (np.asarray(a, dtype=np.float64) for a in lin)
    L = h.shape[-1]
    return (np.ascontiguousarray(np.broadcast_to(Gu, (M, N, L, m))), np.ascontiguousarray(np.broadcast_to(Gx, (M, N, L, n))),
            np.ascontiguousarray(np.broadcast_to(h, (M, N, L))))


def augment(lin, f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref, reg_x, x_l, x_u):
    """The sub-problem with the row values as extra states (see the module docstring).  Returns the augmented
    ``f, fx, fu, x0, X_prev, Q, X_ref, x_l, x_u``; ``U_prev, R, U_ref`` and the control bounds are unchanged."""
    Gu, Gx, h = lin
    M, N, n = f.shape
    m = fu.shape[-1]
    L = h.shape[-1]
    na = n + L
    fa = np.concatenate([f, np.einsum("ijka,ija->ijk", Gu, U_prev) + np.einsum("ijkr,ijr->ijk", Gx, f)], -1)
    fxa = np.zeros((M, N, na, na))
    fxa[..., :n, :n] = fx
    fxa[..., n:, :n] = Gx @ fx
    fua = np.concatenate([fu, Gu + Gx @ fu], -2)
    Qa = np.zeros((M, N, na, na))
    Qa[..., :n, :n] = Q
    Qa[..., np.arange(n, na), np.arange(n, na)] = -float(reg_x)  # cancels the proximal weight added to every state
    pad = lambda a, v: np.concatenate([a, np.full(a.shape[:-1] + (L,), v)], -1)
    absent = lambda z: z is None or np.size(z) == 0
    xl = pad(np.full((M, N, n), -np.inf) if absent(x_l) else np.asarray(x_l, np.float64).reshape(M, N, n), -np.inf)
    xu = np.concatenate([np.full((M, N, n), np.inf) if absent(x_u) else np.asarray(x_u, np.float64).reshape(M, N, n), h], -1)
    return fa, fxa, fua, pad(x0, 0.0), pad(X_prev, 0.0), Qa, pad(X_ref, 0.0), xl, xu

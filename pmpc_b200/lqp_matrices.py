"""Debugging export of the convex sub-problem in canonical form.

The reference exposes ``lqp_generate_problem_matrices`` (``pmpc/scp_mpc.py:66-75`` ->
``PMPC.jl/src/main.jl:374-409``): the matrices ``P, q, A, b, G, l, u`` of

    min 1/2 z'Pz + q'z   s.t.  A z = b,  l <= G z <= u

that its back end hands to OSQP (``lqp_repr_Pq`` / ``lqp_repr_Ab`` / ``lqp_repr_Gla``, ``lqp_utils.jl:2-393``).  The
device solver never forms them — it works on the stage structure directly — so this module assembles them on the host, for
users who inspect a sub-problem or feed it to a solver of their own.  Block by block (one particle and stage at a time), SciPy
sparse, the reference's variable order

    z = [ shared controls u_1..u_Nc ; free controls of particle 1, 2, ... ; states of particle 1, 2, ... ]

(``README.md:232-239``), equality rows particle-major / stage-minor, box rows controls first (shared block: particle 1's
bounds, ``lqp_utils.jl:329-330``) then states.  Arrays are in the Python layout of ``pmpc.solve``: vectors ``(M, N, d)``,
matrices ``(M, N, rows, cols)``; a leading ``M`` may be omitted as in the reference's wrapper.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.sparse as sp


def _nd(a, nd):
    a = np.asarray(a, dtype=np.float64)
    return a.reshape((1,) * (nd - a.ndim) + a.shape) if a.ndim < nd else a


def _per_particle(v, M, default):
    if v is None or np.any(np.isnan(np.asarray(v, dtype=np.float64))):  # NaN sentinel = absent (c_interface.jl:56-70)
        return np.full(M, float(default))
    return np.broadcast_to(np.asarray(v, dtype=np.float64), (M,)).copy()


class _Coo:
    def __init__(self):
        self.r, self.c, self.v = [], [], []

    def block(self, r0, c0, B):
        B = np.atleast_2d(B)
        rr, cc = np.nonzero(B)
        self.r.append(r0 + rr), self.c.append(c0 + cc), self.v.append(B[rr, cc])

    def matrix(self, shape):
        if not self.r:
            return sp.csc_matrix(shape)
        return sp.coo_matrix((np.concatenate(self.v), (np.concatenate(self.r), np.concatenate(self.c))), shape=shape).tocsc()


def lqp_generate_problem_matrices(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx=None, ux=None, lu=None, uu=None,
                                  reg_x=0.0, reg_u=0.0, slew_reg=0.0, slew_reg0=0.0, slew_um1=None, Nc: int = -1):
    """``(P, q, A, b, G, l, u)`` of one SCP iteration's sub-problem (defaults as ``make_probs``, ``main.jl:15-27``:
    no regularisation, no slew penalty, no bounds, full consensus).  ``x0`` only fixes the shapes: the linearisation of stage 1
    is taken at ``x0`` itself, so it does not appear in the matrices (``lqp_utils.jl:288-296``)."""
    f, X_prev, U_prev, X_ref, U_ref = (_nd(a, 3) for a in (f, X_prev, U_prev, X_ref, U_ref))
    fx, fu, Q, R = (_nd(a, 4) for a in (fx, fu, Q, R))
    M, N, n = f.shape
    m = fu.shape[-1]
    Nc = N if int(Nc) < 0 else int(Nc)  # main.jl:127-128
    if not 0 <= Nc <= N:
        raise ValueError("Nc must lie in [0, N]")
    Nf = N - Nc
    rx, ru = _per_particle(reg_x, M, 0.0), _per_particle(reg_u, M, 0.0)
    s = _per_particle(slew_reg, M, 0.0)
    no_um1 = slew_um1 is None or np.any(np.isnan(np.asarray(slew_um1, dtype=np.float64)))
    s0 = np.zeros(M) if no_um1 else _per_particle(slew_reg0, M, 0.0)
    um1 = np.zeros((M, m)) if no_um1 else np.broadcast_to(np.asarray(slew_um1, dtype=np.float64).reshape(-1, m), (M, m))

    def ucol(i, j):  # first column of u_{i,j}
        return j * m if j < Nc else Nc * m + i * Nf * m + (j - Nc) * m

    x_base = Nc * m + M * Nf * m

    def xcol(i, j):
        return x_base + i * N * n + j * n

    nz = x_base + M * N * n
    Pm, Am = _Coo(), _Coo()
    q, b = np.zeros(nz), np.zeros(M * N * n)
    Im, In = np.eye(m), np.eye(n)
    for i in range(M):
        for j in range(N):
            cu, cx = ucol(i, j), xcol(i, j)
            # cost: control block with its proximal weight and the slew diagonal (j = 1: s0 + s; j = N: s; else 2 s)
            w = ru[i] + (s0[i] + s[i] if j == 0 else (s[i] if j == N - 1 else 2.0 * s[i]))
            Pm.block(cu, cu, R[i, j] + w * Im)
            if j + 1 < N and s[i] != 0.0:
                cn = ucol(i, j + 1)
                Pm.block(cu, cn, -s[i] * Im)
                Pm.block(cn, cu, -s[i] * Im)
            q[cu:cu + m] += -ru[i] * U_prev[i, j] - R[i, j] @ U_ref[i, j]
            if j == 0 and Nc >= 1:
                q[cu:cu + m] += -s0[i] * um1[i]  # (with Nc = 0 the reference drops this term: lqp_utils.jl:165 vs :187)
            Pm.block(cx, cx, Q[i, j] + rx[i] * In)
            q[cx:cx + n] = -rx[i] * X_prev[i, j] - Q[i, j] @ X_ref[i, j]
            # dynamics row block (i, j):  fu u_j - x_j + fx x_{j-1} = -f + fu ubar_j + fx xbar_{j-1}
            r0 = (i * N + j) * n
            Am.block(r0, cu, fu[i, j])
            Am.block(r0, cx, -In)
            b[r0:r0 + n] = -f[i, j] + fu[i, j] @ U_prev[i, j]
            if j >= 1:
                Am.block(r0, xcol(i, j - 1), fx[i, j])
                b[r0:r0 + n] += fx[i, j] @ X_prev[i, j - 1]
    P = Pm.matrix((nz, nz))
    A = Am.matrix((M * N * n, nz))

    def present(lo, hi):
        return lo is not None and hi is not None and np.size(lo) and np.size(hi) and not (
            np.any(np.isnan(np.asarray(lo, dtype=np.float64))) or np.any(np.isnan(np.asarray(hi, dtype=np.float64))))

    cols, lo_all, hi_all = [], [], []
    if present(lu, uu):
        lu_, uu_ = _nd(lu, 3), _nd(uu, 3)
        for j in range(Nc):
            cols.append(ucol(0, j) + np.arange(m)), lo_all.append(lu_[0, j]), hi_all.append(uu_[0, j])
        for i in range(M):
            for j in range(Nc, N):
                cols.append(ucol(i, j) + np.arange(m)), lo_all.append(lu_[i, j]), hi_all.append(uu_[i, j])
    if present(lx, ux):
        lx_, ux_ = _nd(lx, 3), _nd(ux, 3)
        for i in range(M):
            for j in range(N):
                cols.append(xcol(i, j) + np.arange(n)), lo_all.append(lx_[i, j]), hi_all.append(ux_[i, j])
    if cols:
        cc = np.concatenate(cols)
        G = sp.coo_matrix((np.ones(cc.size), (np.arange(cc.size), cc)), shape=(cc.size, nz)).tocsc()
        l, u = np.concatenate(lo_all), np.concatenate(hi_all)
    else:
        G, l, u = sp.csc_matrix((0, nz)), np.zeros(0), np.zeros(0)
    return P, q, A, b, G, l, u

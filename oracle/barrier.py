"""Log-barrier ("smoothed constraints") variant of the convex sub-problem.  TEST INFRASTRUCTURE ONLY.

Why it exists: the only outputs of the reference's REAL solver stack (PMPC.jl + ECOS) that the reference repository
holds for fully specified inputs are the iteration tables printed in its notebooks (`examples/gpu_solver.ipynb` cell 18,
`tests/root_testing.ipynb` cells 4 and 11), and those runs pass ``smooth_alpha``, which makes the cone route replace every
box row ``l <= g'z <= u`` by log-barrier terms (`PMPC.jl/src/main.jl:240-262`: ``haskey(settings, :smooth_alpha)`` selects
``smooth_cstr = "logbarrier"``; `cone_utils.jl:172-200` ``make_logbarrier_constraint``: an epigraph variable
``t >= -log(alpha * slack) / alpha`` per row, all of them summed into the objective with weight one).  Restating that
variant on top of THIS oracle's matrices (``oracle.lqp.build_lqp``: P, q, A, b, G, l, u) lets the printed tables pin the
oracle's costs, proximal weights, slew penalty, dynamics rows, consensus layout and SCP bookkeeping against numbers the
reference itself produced (tests/test_reference_notebook_traces.py); the hard-box QP the CUDA path solves differs from
it only in how the box rows are treated.

The cone route minimises ``(1 + e) sum_i y_i + (1 - e) k t`` subject to ``cost_i(z) <= y_i + t, y_i >= 0`` with
``e = COST_ANCHOR_EPS = 1e-3`` and ``k = M`` (`main.jl:219-233`, `cone_utils.jl:64-96`).  Minimising over ``y, t`` for
fixed ``z`` leaves ``(1 + e) sum_i cost_i - 2 e M min_i cost_i``: for one particle ``(1 - e) cost``, for several the sum
of costs with the cheapest particle's weight lowered by ``2 e M``.  ``cost_i = 1/2 z_i' P_i z_i + q_i' z_i`` is the same
quadratic the QP route sums (`qp_utils.jl:59-150` against `lqp_utils.jl:2-216`).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .lqp import build_lqp, split_vars

COST_ANCHOR_EPS = 1e-3  # main.jl:222


def barrier_lqp_solve(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, smooth_alpha, anchor_eps=COST_ANCHOR_EPS, tol=1e-11,
                      max_newton=200, **kw):
    """Minimiser of the cone route's smoothed sub-problem (see the module docstring) by an equality-constrained Newton
    method (sparse KKT solves) started strictly inside the box.  Same arguments as ``oracle.lqp.lqp_solve`` plus
    ``smooth_alpha``; ``anchor_eps = 0`` gives the plain sum of the particle costs under the same barrier (the objective
    of the QP route).  Returns ``X (M,N,n), U (M,N,m), info``."""
    x0, f, fx, fu = (np.asarray(a, float) for a in (x0, f, fx, fu))
    X_prev, U_prev, Q, R, X_ref, U_ref = (np.asarray(a, float) for a in (X_prev, U_prev, Q, R, X_ref, U_ref))
    prob = build_lqp(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, **kw)
    M, alpha, e, nz = prob.M, float(smooth_alpha), float(anchor_eps), prob.nz
    A, b, G, lo, hi = prob.A.tocsr(), prob.b, prob.G.tocsr(), prob.l, prob.u
    # cost_i as a quadratic in the whole variable vector: particle i's own sub-problem (same builder, one particle, its
    # order [u (N m); x (N n)]) scattered to the consensus layout, in which the first Nc controls are shared
    per_particle = ("lx", "ux", "lu", "uu", "slew_um1", "slew_reg", "slew_reg0", "reg_x", "reg_u")
    W, w0 = [], []
    for i in range(M):
        kwi = {k: v for k, v in kw.items() if k != "Nc"}
        for k in per_particle:
            if kwi.get(k) is not None and np.ndim(kwi[k]) > 0:
                kwi[k] = np.asarray(kwi[k])[i:i + 1]
        one = build_lqp(x0[i:i + 1], f[i:i + 1], fx[i:i + 1], fu[i:i + 1], X_prev[i:i + 1], U_prev[i:i + 1], Q[i:i + 1],
                        R[i:i + 1], X_ref[i:i + 1], U_ref[i:i + 1], Nc=0, **kwi)
        idx = np.concatenate([prob.iu[i].ravel(), prob.ix[i].ravel()])
        S = sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, nz))
        W.append((S.T @ one.P @ S).tocsr())
        w0.append(S.T @ one.q)

    def costs(z):
        return np.array([0.5 * z @ (Wi @ z) + qi @ z for Wi, qi in zip(W, w0)])

    def weights(c):  # (1 + e) sum_i cost_i - 2 e M min_i cost_i
        w = np.full(M, 1.0 + e)
        w[int(np.argmin(c))] -= 2.0 * e * M
        return w

    def value(z):
        su, sl = hi - G @ z, G @ z - lo
        if (su <= 0).any() or (sl <= 0).any():
            return np.inf
        c = costs(z)
        return float(weights(c) @ c) - (np.log(alpha * su).sum() + np.log(alpha * sl).sum()) / alpha

    # strictly feasible start: box midpoints on the rows' variables; the equality residual goes with the first full step
    z = np.zeros(nz)
    if G.shape[0]:
        Gc = G.tocoo()
        z[Gc.col] = (0.5 * (lo + hi))[Gc.row] / Gc.data
    nA = A.shape[0]
    its, val = 0, value(z)
    for its in range(1, max_newton + 1):
        su, sl = hi - G @ z, G @ z - lo
        w = weights(costs(z))
        grad = sum(wi * (Wi @ z + qi) for wi, Wi, qi in zip(w, W, w0)) + G.T @ (1.0 / su - 1.0 / sl) / alpha
        H = sum(wi * Wi for wi, Wi in zip(w, W)) + G.T @ sp.diags((1.0 / su ** 2 + 1.0 / sl ** 2) / alpha) @ G
        K = sp.bmat([[H, A.T], [A, None]], format="csc")
        r_eq = np.abs(A @ z - b).max() if nA else 0.0
        dz = spla.spsolve(K, -np.concatenate([grad, A @ z - b]))[:nz]
        if np.abs(dz).max() < tol and r_eq < 1e-10:
            break
        t = 1.0
        while True:  # stay strictly inside the box; once the equalities hold, also require descent
            zn = z + t * dz
            vn = value(zn)
            if np.isfinite(vn) and (r_eq > 1e-9 or vn <= val + 1e-4 * t * (grad @ dz) + 1e-13 * abs(val)):
                break
            t *= 0.5
            if t < 1e-12:
                raise RuntimeError("barrier Newton: line search failed")
        z, val = zn, vn
    X, U = split_vars(prob, z)
    return X, U, dict(newton_iterations=its, objective=val)

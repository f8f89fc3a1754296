"""Seeded problem generators shared by the CPU and GPU tests and by bench.py.

* ``double_integrator`` re-creates the deterministic INPUT fixture the reference sweeps through
  its C ABI (`PMPC.jl/src/c_precompile.jl:6-49`, `PMPC.jl/pmpcjl/test.py:171-226`); the reference
  stores no expected outputs for it, the oracle supplies them.
* ``config_c1 .. config_c5`` are the five BASELINE.json configs as specified in SURVEY.md 8(d).
"""
import numpy as np


def rand_problem(rng, M, N, n, m, slew=True, box=0.1, xbox=None, per_particle_slew=True):
    spd = lambda k: (lambda a: a @ a.T / k + 0.5 * np.eye(k))(rng.standard_normal((k, k)))
    Q = np.stack([[spd(n) for _ in range(N)] for _ in range(M)])
    R = np.stack([[spd(m) for _ in range(N)] for _ in range(M)])
    fx = np.eye(n) + 0.2 * rng.standard_normal((M, N, n, n))
    fu = rng.standard_normal((M, N, n, m))
    f = rng.standard_normal((M, N, n))
    x0 = rng.standard_normal((M, n))
    X_prev, U_prev = rng.standard_normal((M, N, n)), 0.3 * rng.standard_normal((M, N, m))
    X_ref, U_ref = rng.standard_normal((M, N, n)), 0.3 * rng.standard_normal((M, N, m))
    kw = dict(reg_x=1.0, reg_u=0.1)
    if slew:
        s = rng.uniform(0.5, 1.5, M) if per_particle_slew else np.full(M, 0.8)
        kw.update(slew_reg=s, slew_reg0=rng.uniform(0.5, 1.5, M), slew_um1=rng.standard_normal((M, m)))
    if box:
        kw.update(lu=-box * np.ones((M, N, m)), uu=box * np.ones((M, N, m)))
    if xbox:
        kw.update(lx=-xbox * np.ones((M, N, n)), ux=xbox * np.ones((M, N, n)))
    return (x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref), kw


def double_integrator(u_limit=0.4, M=1):
    """c_precompile.jl:6-49 / pmpcjl/test.py:171-226 in Python layout. Returns (args, kw, Nc)."""
    N, n, m, Nc = 30, 2, 1, 3
    x0 = np.tile(np.array([5.0, 5.0]), (M, 1))
    A = np.array([[1.0, 0.1], [0.0, 1.0]])
    fx = np.tile(A, (M, N, 1, 1))
    fu = np.tile(np.array([[0.0], [1.0]]), (M, N, 1, 1))
    f = np.zeros((M, N, n))
    f[:, 0] = x0 @ A.T
    X_prev, U_prev = np.zeros((M, N, n)), np.zeros((M, N, m))
    X_ref, U_ref = np.zeros((M, N, n)), np.zeros((M, N, m))
    Q, R = np.tile(np.eye(n), (M, N, 1, 1)), np.tile(np.eye(m), (M, N, 1, 1))
    kw = dict(reg_x=1.0, reg_u=0.1, lu=-u_limit * np.ones((M, N, m)), uu=u_limit * np.ones((M, N, m)),
              lx=-20.0 * np.ones((M, N, n)), ux=20.0 * np.ones((M, N, n)), slew_reg=np.ones(M),
              slew_reg0=np.zeros(M), slew_um1=np.zeros((M, m)))
    return (x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref), kw, Nc


def _dubins_base(M, N):
    n, m = 4, 2
    return dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(1e-2 * np.eye(m), (M, N, 1, 1)), x0=np.ones((M, n)),
                X_ref=np.zeros((M, N, n)), U_ref=np.zeros((M, N, m)), X_prev=np.zeros((M, N, n)),
                U_prev=np.zeros((M, N, m)), u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)), reg_x=1.0, reg_u=1e-2)


def config_c1():
    """C1 (seed 0): Dubins p=(1,1,0.3), N=20, M=1, exactly 10 iterations (res_tol=-1)."""
    d = _dubins_base(1, 20)
    d.update(max_it=10, res_tol=-1.0, solver_settings=dict(solver="osqp"))
    return d, np.array([[1.0, 1.0, 0.3]])


def config_c2(M=128, N=50):
    """C2 (seed 1): perturbed-dynamics particles v,w ~ U[0.7,1], Nc=1, res_tol=1e-5, max_it=100."""
    rng = np.random.default_rng(1)
    p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)
    d = _dubins_base(M, N)
    d.update(max_it=100, res_tol=1e-5, solver_settings=dict(solver="osqp", Nc=1))
    return d, p


def config_c3(B=16384, M=16, N=30):
    """C3 (seed 2): B independent problems, x0 ~ N(1,0.1^2), X_ref = 0.1 N(0,1), Nc=1."""
    rng = np.random.default_rng(2)
    n, m = 4, 2
    p = np.stack([rng.uniform(0.7, 1.0, (B, M)), rng.uniform(0.7, 1.0, (B, M)), np.full((B, M), 0.3)], -1)
    x0 = np.repeat(1.0 + 0.1 * rng.standard_normal((B, 1, n)), M, axis=1)
    X_ref = np.repeat(0.1 * rng.standard_normal((B, 1, N, n)), M, axis=1)
    d = dict(Q=np.broadcast_to(np.eye(n), (B, M, N, n, n)), R=np.broadcast_to(1e-2 * np.eye(m), (B, M, N, m, m)), x0=x0,
             X_ref=X_ref, U_ref=np.zeros((B, M, N, m)), X_prev=np.zeros((B, M, N, n)), U_prev=np.zeros((B, M, N, m)),
             u_l=-np.ones((B, M, N, m)), u_u=np.ones((B, M, N, m)), reg_x=1.0, reg_u=1e-2, max_it=100, res_tol=1e-5,
             solver_settings=dict(solver="osqp", Nc=1))
    return d, p


def config_ltv(seed, M, N, n, m, Nc, xbox=True):
    """C4 / C5 generator: A_j = I + 0.05 G/sqrt(n), B_j = 0.1 G, f exactly linear."""
    rng = np.random.default_rng(seed)
    A = np.eye(n) + 0.05 * rng.standard_normal((M, N, n, n)) / np.sqrt(n)
    Bm = 0.1 * rng.standard_normal((M, N, n, m))
    d = dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(0.1 * np.eye(m), (M, N, 1, 1)),
             x0=np.tile(rng.standard_normal(n), (M, 1)), X_ref=np.zeros((M, N, n)), U_ref=np.zeros((M, N, m)),
             X_prev=np.zeros((M, N, n)), U_prev=np.zeros((M, N, m)), u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)),
             reg_x=1.0, reg_u=1e-2, slew_rate=1.0, max_it=20, res_tol=1e-7, solver_settings=dict(solver="osqp", Nc=Nc))
    if xbox:
        d.update(x_l=-20.0 * np.ones((M, N, n)), x_u=20.0 * np.ones((M, N, n)))
    return d, A, Bm


def config_c4(M=64, N=100, n_lin=0):
    """C4 (seed 3): 12-state / 4-input LTV, slew, control and state boxes, Nc = 1; with ``n_lin`` > 0 also SURVEY 8(d)'s
    ``n_lin`` random linear rows g_k'[u; x] <= h_k per (particle, stage) (BASELINE: "box/linear constraints"), returned in
    ``solver_settings["lin_cstrs"] = (Gu, Gx, h)``.  h_k = 0.5 + |g_k'[0; x~]| with x~ the uncontrolled response, so the
    rows hold with margin at u = 0 (the problem is feasible) and bind once the optimal controls push against them."""
    d, A, Bm = config_ltv(3, M, N, 12, 4, 1, xbox=True)
    if n_lin > 0:
        rng = np.random.default_rng(33)
        n, m = 12, 4
        Gu = rng.standard_normal((M, N, n_lin, m)) / np.sqrt(n + m)
        Gx = rng.standard_normal((M, N, n_lin, n)) / np.sqrt(n + m)
        xt = np.empty((M, N, n))
        x = d["x0"].copy()
        for j in range(N):
            x = np.einsum("mrt,mt->mr", A[:, j], x)
            xt[:, j] = x
        h = 0.5 + np.abs(np.einsum("mjkr,mjr->mjk", Gx, xt))
        d["solver_settings"] = dict(d["solver_settings"], lin_cstrs=(Gu, Gx, h))
    return d, A, Bm


def config_c5(M=1024, N=200):
    return config_ltv(4, M, N, 32, 8, 5, xbox=False)


def hooks_lin_cost_fn(X_prev, U_prev, problems):
    """The `lin_cost_fn` of tests/golden/scp_lin_cost.npz: gradient of the smooth extra cost
    0.05 sum(sin(x)) + 1e-3 sum(u^3) / 3 at the previous iterate (reference signature, pmpc/scp_mpc.py:171-175)."""
    return 0.05 * np.cos(X_prev), 1e-3 * U_prev ** 2

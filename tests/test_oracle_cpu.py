"""CPU tests of the oracle (test infrastructure) against the reference-generated golden files and
against itself (solver-independent KKT certificate).  No GPU needed."""
import os

import numpy as np
import pytest

from fixtures import double_integrator, rand_problem
from oracle import dubins, lqp, riccati_admm, scp
from oracle.osqp_port import osqp_solve

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _direct_cost(args, kw, X, U, Nc):
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    M, N, n = f.shape
    m = fu.shape[-1]
    s, s0 = kw.get("slew_reg", np.zeros(M)), kw.get("slew_reg0", np.zeros(M))
    um1 = kw.get("slew_um1", np.zeros((M, m)))
    J = 0.0
    for i in range(M):
        for j in range(N):
            dx, du = X[i, j] - Xr[i, j], U[i, j] - Ur[i, j]
            J += 0.5 * dx @ Q[i, j] @ dx + 0.5 * kw["reg_x"] * np.sum((X[i, j] - Xp[i, j]) ** 2)
            J += 0.5 * du @ R[i, j] @ du + 0.5 * kw["reg_u"] * np.sum((U[i, j] - Up[i, j]) ** 2)
            if j >= 1:
                J += 0.5 * s[i] * np.sum((U[i, j] - U[i, j - 1]) ** 2)
        J += 0.5 * s0[i] * np.sum((U[i, 0] - (um1[i] if Nc >= 1 else 0.0)) ** 2)  # Nc == 0 quirk (App. B.6)
        if N == 1:
            J += 0.5 * s[i] * np.sum(U[i, 0] ** 2)  # lqp_utils.jl:31-37
    return J


@pytest.mark.parametrize("dims", [(3, 6, 4, 2, 2), (1, 5, 2, 1, 0), (4, 7, 3, 2, -1), (2, 4, 3, 2, 1), (2, 1, 2, 1, 1)])
def test_assembled_qp_is_the_sum_of_particle_costs(dims):
    """P, q, A, b built like lqp_repr_Pq/Ab reproduce sum_i J_i and the linearised dynamics."""
    M, N, n, m, Nc = dims
    rng = np.random.default_rng(sum(dims))
    args, kw = rand_problem(rng, M, N, n, m)
    prob = lqp.build_lqp(*args, Nc=Nc, **kw)
    Nce = N if Nc < 0 else Nc
    z1, z2 = rng.standard_normal(prob.nz), rng.standard_normal(prob.nz)
    cost = lambda z: 0.5 * z @ (prob.P @ z) + prob.q @ z
    (X1, U1), (X2, U2) = lqp.split_vars(prob, z1), lqp.split_vars(prob, z2)
    d_qp = cost(z1) - cost(z2)
    d_dir = _direct_cost(args, kw, X1, U1, Nce) - _direct_cost(args, kw, X2, U2, Nce)
    assert abs(d_qp - d_dir) <= 1e-10 * max(1.0, abs(d_dir))
    x0, f, fx, fu, Xp, Up = args[:6]
    pred = f + np.einsum("ijrt,ijt->ijr", fu, U1 - Up)
    pred[:, 1:] += np.einsum("ijrt,ijt->ijr", fx[:, 1:], X1[:, :-1] - Xp[:, :-1])
    assert np.abs((prob.A @ z1 - prob.b) - (pred - X1).ravel()).max() < 1e-12
    # consensus: the first Nc controls are literally shared
    assert np.all(U1[:, :Nce] == U1[0:1, :Nce])


def test_exact_solver_certificate_and_osqp_port_agree():
    rng = np.random.default_rng(5)
    args, kw = rand_problem(rng, 4, 8, 4, 2, box=0.15)
    X, U, info = lqp.lqp_solve(*args, Nc=2, **kw)
    cert = info["cert"]
    assert cert.worst < 1e-9
    assert np.sum(np.abs(U) > 0.15 - 1e-12) > 5  # constraints really are active
    res = osqp_solve(info["prob"])  # reference settings: eps 1e-3
    assert res.status == "solved"
    assert np.abs(res.x - info["z"]).max() < 5e-2
    res = osqp_solve(info["prob"], eps_abs=1e-9, eps_rel=1e-9, max_iter=100000)
    assert np.abs(res.x - info["z"]).max() < 1e-6


def test_sentinel_sweep_double_integrator():
    """The reference sweeps present/absent combinations of every optional argument through its C
    ABI (c_precompile.jl:52-100); same sweep through the oracle: each result must certify."""
    args, kw, Nc = double_integrator(u_limit=0.4)
    for ub in (False, True):
        for xb in (False, True):
            for s0 in (False, True):
                for s in (False, True):
                    k = dict(reg_x=kw["reg_x"], reg_u=kw["reg_u"])
                    if ub:
                        k.update(lu=kw["lu"], uu=kw["uu"])
                    if xb:
                        k.update(lx=kw["lx"], ux=kw["ux"])
                    k["slew_reg"] = kw["slew_reg"] if s else np.full(1, np.nan)
                    k["slew_reg0"] = kw["slew_reg0"] if s0 else np.full(1, np.nan)
                    k["slew_um1"] = kw["slew_um1"] if s0 else np.full((1, 1), np.nan)
                    X, U, info = lqp.lqp_solve(*args, Nc=Nc, **k)
                    assert info["cert"].worst < 1e-9
                    if ub:
                        assert np.abs(U).max() <= 0.4 + 1e-9


def test_dubins_matches_reference_autograd_where_well_conditioned():
    """oracle/dubins.py vs the reference's torch-autograd values (golden): 1e-12 relative for
    |u| >= 0.05; for tiny |u| the REFERENCE's own values are noise: the stable form
    must stay within the reference's conditioning bound ~1e-16/|a|^3."""
    g = np.load(os.path.join(GOLD, "dubins_lin.npz"))
    rel = lambda a, b: np.abs(a - b).max() / max(1.0, np.abs(b).max())
    f, fx, fu = dubins.car_f_fx_fu(g["generic_X"], g["generic_U"], g["generic_p"])
    assert rel(f, g["generic_f"]) < 1e-12 and rel(fx, g["generic_fx"]) < 1e-12 and rel(fu, g["generic_fu"]) < 1e-12
    fn, fxn, fun = dubins.car_f_fx_fu_naive(g["generic_X"], g["generic_U"], g["generic_p"])
    assert rel(fn, g["generic_f"]) < 1e-12 and rel(fun, g["generic_fu"]) < 1e-12
    for k, tol_x, tol_u in (("small_u", 1e-8, 1e-3), ("zero_u", 1e-8, 1e-1)):
        f, fx, fu = dubins.car_f_fx_fu(g[k + "_X"], g[k + "_U"], g[k + "_p"])
        assert rel(f, g[k + "_f"]) < tol_x and rel(fx, g[k + "_fx"]) < tol_x and rel(fu, g[k + "_fu"]) < tol_u


def test_dubins_stable_form_against_high_precision():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(3)
    for u2 in (0.7, 1e-3, 1e-6, 0.0):
        x, u, p = rng.standard_normal(4), np.array([0.3, u2]), np.array([0.9, 0.8, 0.3])
        f, fx, fu = dubins.car_f_fx_fu(x, u, p)
        vs, ws, T = [mp.mpf(float(v)) for v in p]
        b, a = vs * mp.mpf(float(u[0])) + mp.mpf(1e-6), ws * (-mp.mpf(float(u[1])))
        a = a + (mp.mpf(1e-6) if a >= 0 else -mp.mpf(1e-6))
        x0, y0, v0, th = [mp.mpf(float(v)) for v in x]
        ph = T * a + th
        xp1 = x0 + (a * mp.sin(ph) * v0 + T * b * a * mp.sin(ph) + b * mp.cos(ph)) / a**2 - (mp.sin(th) * a * v0 + mp.cos(th) * b) / a**2
        xp2 = y0 - (a * mp.cos(ph) * v0 - b * mp.sin(ph) + T * b * a * mp.cos(ph)) / a**2 + (mp.cos(th) * a * v0 - mp.sin(th) * b) / a**2
        assert abs(float(xp1) - f[0]) < 1e-14 and abs(float(xp2) - f[1]) < 1e-14


@pytest.mark.parametrize("name,kw", [
    ("scp_c1", dict(max_it=10, res_tol=-1.0)),
    ("scp_consensus", dict(max_it=40, res_tol=1e-7, solver_settings=dict(Nc=1))),
    ("scp_slew", dict(max_it=60, res_tol=1e-7, slew_rate=0.5, solver_settings=dict(Nc=2, slew_reg=0.3))),
])
def test_scp_loop_replays_reference_loop(name, kw):
    """oracle.scp.scp_solve fed with the linearisations the REFERENCE loop consumed must reproduce
    the reference loop's X, U and (obj, resid) history (golden generated by make_golden.py)."""
    d = np.load(os.path.join(GOLD, name + ".npz"))
    it = iter(range(len(d["lin_f"])))

    def replay(X, U):
        k = next(it)
        return d["lin_f"][k], d["lin_fx"][k], d["lin_fu"][k]

    extra = {}
    if name == "scp_slew":
        extra = dict(u0_slew=0.2 * np.ones((3, 2)), x_l=d["x_l"], x_u=d["x_u"])
    X, U, data = scp.scp_solve(replay, d["Q"], d["R"], d["x0"], X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"],
                               U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], reg_x=1.0, reg_u=1e-2, **kw, **extra)
    hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
    assert len(hist) == len(d["hist"])
    assert np.abs(X - d["X"]).max() < 1e-9 and np.abs(U - d["U"]).max() < 1e-9
    assert np.abs(hist - d["hist"]).max() < 1e-9


@pytest.mark.parametrize("dims", [(3, 6, 4, 2, 2, True, 0.1), (1, 5, 2, 1, 0, True, 0.1), (4, 7, 3, 2, -1, True, 0.2),
                                  (5, 8, 4, 2, 3, False, 0.3), (3, 6, 4, 2, 0, True, 0.1), (5, 8, 4, 2, 1, True, None)])
def test_port_of_device_algorithm_matches_exact_oracle(dims):
    """The NumPy port of the Riccati-ADMM + polish algorithm (what the kernels implement) reaches
    the exact QP solution."""
    M, N, n, m, Nc, slew, box = dims
    rng = np.random.default_rng(7 + sum(dims[:5]))
    args, kw = rand_problem(rng, M, N, n, m, slew, box)
    Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
    X, U, info = riccati_admm.lqp_solve(*args, Nc=Nc, **kw)
    assert np.abs(X - Xo).max() < 1e-8 and np.abs(U - Uo).max() < 1e-8
    X, U, info = riccati_admm.lqp_solve(*args, Nc=Nc, settings=riccati_admm.Settings(polish=False), **kw)
    assert np.abs(X - Xo).max() < 1e-6 and np.abs(U - Uo).max() < 1e-6


# ---- the plain-C oracle (oracle/c/pmpc_oracle.c: the compiled CPU baseline) pinned against the NumPy oracle
# and the reference-generated golden files -------------------------------------------------------------------
from oracle import cport  # noqa: E402

TIGHT = dict(eps_abs=1e-10, eps_rel=1e-10, max_iter=200000)


@pytest.mark.parametrize("dims", [(3, 6, 4, 2, 2, True, 0.1), (1, 5, 2, 1, 0, True, 0.1), (4, 7, 3, 2, -1, True, 0.2),
                                  (2, 4, 3, 2, 1, False, 0.1), (2, 1, 2, 1, 1, True, 0.1), (5, 8, 4, 2, 1, True, None)])
def test_c_oracle_qp_matches_exact_oracle(dims):
    """QP assembly (lqp_utils.jl restated in C) + OSQP restated in C, run to tight tolerance, against the exact KKT
    solve of oracle/lqp.py on the same seeded inputs (slew, u_{-1} term, every Nc case, boxes present / absent)."""
    M, N, n, m, Nc, slew, box = dims
    rng = np.random.default_rng(7 + sum(dims[:5]))
    args, kw = rand_problem(rng, M, N, n, m, slew, box, None)
    Xo, Uo, info = lqp.lqp_solve(*args, Nc=Nc, **kw)
    Xc, Uc, inf = cport.lqp_solve(*args, Nc=Nc, osqp_settings=TIGHT, **kw)
    assert inf.status == 1
    assert np.abs(Xc - Xo).max() < 1e-7 and np.abs(Uc - Uo).max() < 1e-7


def test_c_oracle_sentinels_and_default_tolerance():
    """NaN sentinels (c_interface.jl:56-70) on the reference's double-integrator fixture; at OSQP's default 1e-3
    tolerances (what the reference runs with) the C port stays within OSQP accuracy of the exact minimiser."""
    args, kw, Nc = double_integrator()
    k = dict(kw)
    k["lx"] = np.full_like(kw["lx"], np.nan)            # state box disabled by sentinel
    k["slew_reg0"] = np.full(1, np.nan)                 # u_{-1} term disabled by sentinel
    Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **k)
    Xc, Uc, inf = cport.lqp_solve(*args, Nc=Nc, osqp_settings=TIGHT, **k)
    assert inf.status == 1 and np.abs(Xc - Xo).max() < 1e-7 and np.abs(Uc - Uo).max() < 1e-7
    Xd, Ud, infd = cport.lqp_solve(*args, Nc=Nc, **k)
    assert infd.status == 1 and infd.iters < inf.iters and np.abs(Ud - Uo).max() < 5e-2


def test_c_oracle_dubins_matches_reference_golden():
    """The C Dubins model against the reference's own torch-autograd values (tests/golden/dubins_lin.npz)."""
    import ctypes as C

    g = np.load(os.path.join(GOLD, "dubins_lin.npz"))
    lib = cport.load()
    X, U, P = g["generic_X"].reshape(-1, 4), g["generic_U"].reshape(-1, 2), np.broadcast_to(g["generic_p"], g["generic_X"].shape[:-1] + (3,)).reshape(-1, 3)
    f, fx, fu = np.empty((len(X), 4)), np.empty((len(X), 16)), np.empty((len(X), 8))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for k in range(len(X)):
        x, u, p = np.ascontiguousarray(X[k]), np.ascontiguousarray(U[k]), np.ascontiguousarray(P[k])
        lib.pmpc_oracle_dubins(dp(x), dp(u), dp(p), C.c_double(1e-6), dp(f[k]), dp(fx[k]), dp(fu[k]))
    rel = lambda a, b: np.abs(a - b.reshape(a.shape)).max() / max(1.0, np.abs(b).max())
    assert rel(f, g["generic_f"]) < 1e-12 and rel(fx, g["generic_fx"]) < 1e-12 and rel(fu, g["generic_fu"]) < 1e-12


@pytest.mark.parametrize("name,kw,tol", [("scp_c1", dict(max_it=10, res_tol=-1.0, Nc=-1), 5e-4),
                                         ("scp_consensus", dict(max_it=40, res_tol=1e-7, Nc=1), 1e-6)])
def test_c_oracle_scp_loop_against_reference_loop_golden(name, kw, tol):
    """The whole C loop (linearise, assemble, OSQP at tight tolerance, bookkeeping) against the trajectories the
    REFERENCE's scp_solve produced here (make_golden.py).  The converged consensus case agrees to 1e-6; the ten forced
    iterations of C1 start at U_prev = 0, where the reference's autograd Jacobian carries ~1e-3 absolute noise
    (SURVEY.md H3), which has not yet contracted away after ten iterations: 5e-4."""
    d = np.load(os.path.join(GOLD, name + ".npz"))
    M = d["Q"].shape[0]
    p = np.broadcast_to(d["p"].reshape(-1, 3), (M, 3))
    X, U, h = cport.scp_dubins(p, d["Q"], d["R"], d["x0"], d["X_ref"], d["U_ref"], d["X_prev"], d["U_prev"], d["u_l"], d["u_u"],
                               osqp_settings=TIGHT, **kw)
    assert len(h) == len(d["hist"])
    assert np.abs(X - d["X"]).max() < tol and np.abs(U - d["U"]).max() < tol
    assert abs(h[-1, 0] - d["hist"][-1, 0]) < tol * max(1.0, abs(d["hist"][-1, 0]))


def test_c_oracle_batch_threads_equal_serial():
    """pthread batch entry (one problem per host thread, the shape of the reference's worker pool) == serial solves."""
    from fixtures import config_c3

    B = 4
    d, p = config_c3(B=B, M=3, N=8)
    st = dict(eps_abs=1e-8, eps_rel=1e-8, max_iter=50000, adaptive_rho_interval=25)  # fixed interval: timing-independent
    Xb, Ub, its = cport.scp_dubins_batch(2, p, d["Q"], d["R"], d["x0"], d["X_ref"], d["U_ref"], d["X_prev"], d["U_prev"],
                                         d["u_l"], d["u_u"], max_it=30, res_tol=1e-6, Nc=1, osqp_settings=st)
    for b in range(B):
        X, U, h = cport.scp_dubins(p[b], d["Q"][b], d["R"][b], d["x0"][b], d["X_ref"][b], d["U_ref"][b], d["X_prev"][b],
                                   d["U_prev"][b], d["u_l"][b], d["u_u"][b], max_it=30, res_tol=1e-6, Nc=1, osqp_settings=st)
        assert its[b] == len(h) and np.array_equal(X, Xb[b]) and np.array_equal(U, Ub[b])


def test_exact_qp_oracle_against_scipy_trust_constr():
    """The exact QP oracle (own primal-dual active-set solve + KKT certificate) cross-checked by an INDEPENDENT third-party
    solver on the same assembled matrices: SciPy's trust-constr interior-point method on  min 1/2 z'Pz + q'z,  A z = b,
    l <= G z <= u  (consensus, slew, u_{-1} term, control and state boxes, general linear rows).  Not the reference's
    solver stack (Julia / OSQP cannot run here), but a solver this repository did not write."""
    from scipy.optimize import Bounds, LinearConstraint, minimize

    from fixtures import rand_problem

    rng = np.random.default_rng(70)
    M, N, n, m, Nc = 2, 5, 3, 2, 1
    args, kw = rand_problem(rng, M, N, n, m, True, 0.6, per_particle_slew=False)
    Xf, Uf, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
    xb = 0.9 * np.abs(Xf[:, N - 2:]).max()
    lx, ux = -1e3 * np.ones((M, N, n)), 1e3 * np.ones((M, N, n))
    lx[:, N - 2:], ux[:, N - 2:] = -xb, xb
    Gu, Gx = rng.standard_normal((M, N, 1, m)), rng.standard_normal((M, N, 1, n))
    val = np.einsum("ijka,ija->ijk", Gu, Uf) + np.einsum("ijkr,ijr->ijk", Gx, Xf)
    h = val + np.where(rng.random(val.shape) < 0.3, -0.05 * (0.2 + np.abs(val)), 1.0)
    X, U, info = lqp.lqp_solve(*args, Nc=Nc, lx=lx, ux=ux, lin_rows=(Gu, Gx, h), **kw)
    prob = info["prob"]
    P, q, A, b, G = prob.P.toarray(), prob.q, prob.A.toarray(), prob.b, prob.G.toarray()
    cons = [LinearConstraint(A, b, b), LinearConstraint(G, prob.l, prob.u)]
    res = minimize(lambda z: 0.5 * z @ P @ z + q @ z, np.zeros(prob.nz), jac=lambda z: P @ z + q, hess=lambda z: P,
                   constraints=cons, method="trust-constr", options=dict(gtol=1e-12, xtol=1e-14, barrier_tol=1e-12, maxiter=3000))
    assert np.abs(A @ res.x - b).max() < 1e-8
    assert abs(res.fun - info["cert"].objective) <= 1e-7 * max(1.0, abs(info["cert"].objective))
    assert np.abs(res.x - info["z"]).max() <= 1e-5 * max(1.0, np.abs(info["z"]).max())
    assert np.sum(np.abs(prob.G @ info["z"] - prob.u) < 1e-9) + np.sum(np.abs(prob.G @ info["z"] - prob.l) < 1e-9) >= 3  # active rows

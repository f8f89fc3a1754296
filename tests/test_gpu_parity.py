"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the reference-generated golden fixtures, and — at BASELINE sizes — through
size-independent properties.  Tolerances (BASELINE.json north_star): linearisation 1e-12 relative,
converged X, U 1e-5 relative, objective 1e-6 relative."""
import os

import numpy as np
import pytest

import pmpc_b200
from fixtures import config_c1, config_c2, config_c3, config_c4, double_integrator, rand_problem
from oracle import dubins, lqp, scp as oscp
from pmpc_b200 import DeviceSolver, DubinsCar, LTV, static_backend

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


def _aff(args, kw, Nc, solver="osqp"):
    """One sub-problem through aff_solve -> lqp_solve -> pmpcjl.lqp_solve -> c_lqp_solve (C ABI)."""
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    ss = dict(solver=solver, Nc=Nc)
    slew_rate = None
    if "slew_reg" in kw and not np.isnan(kw["slew_reg"]).any():
        slew_rate = float(kw["slew_reg"][0])
    u_slew = None
    if "slew_reg0" in kw and not np.isnan(kw["slew_reg0"]).any():
        ss["slew_reg"] = float(kw["slew_reg0"][0])
        u_slew = kw["slew_um1"]
    X, U, _ = static_backend.aff_solve(f, fx, fu, x0, Xp, Up, Q, R, Xr, Ur, kw["reg_x"], kw["reg_u"], slew_rate, u_slew,
                                       kw.get("lx"), kw.get("ux"), kw.get("lu"), kw.get("uu"), ss)
    return X[:, 1:], U


CASES = [
    # M, N, n, m, Nc, slew, ubox, xbox
    (1, 5, 2, 1, 0, True, 0.1, None), (3, 6, 4, 2, 2, True, 0.1, None), (4, 7, 3, 2, -1, True, 0.2, None),
    (2, 4, 3, 2, 1, False, 0.1, None), (2, 1, 2, 1, 1, True, 0.1, None), (5, 8, 4, 2, 1, True, None, None),
    (5, 8, 4, 2, 3, False, 0.3, None), (3, 6, 4, 2, 0, True, 0.1, None), (40, 10, 4, 2, 1, True, 0.3, None),
    (3, 5, 6, 3, 1, True, 0.2, None), (2, 6, 12, 4, 1, True, 0.2, None), (1, 9, 4, 2, 3, True, 0.2, None),
    (33, 5, 4, 2, 2, True, 0.2, None), (2, 4, 7, 5, 1, False, 0.2, None),
]


@pytest.mark.parametrize("case", CASES)
def test_qp_parity_through_c_abi(gpu, case):
    """c_lqp_solve vs the exact oracle (uniform per-problem slew, the only kind the Python path can
    express: static_backend.py:262-267 broadcasts one scalar)."""
    M, N, n, m, Nc, slew, box, xbox = case
    rng = np.random.default_rng(100 + sum(x for x in case[:5]))
    args, kw = rand_problem(rng, M, N, n, m, slew, box, xbox, per_particle_slew=False)
    if slew:
        kw["slew_reg0"] = np.full(M, 0.7)
    Xo, Uo, info = lqp.lqp_solve(*args, Nc=Nc, **kw)
    X, U = _aff(args, kw, Nc)
    assert rel(X, Xo) < 1e-7 and rel(U, Uo) < 1e-7  # target 1e-5; the polish is exact to round-off
    zo = info["z"]
    obj = lambda Xs, Us: (lambda z: 0.5 * z @ (info["prob"].P @ z) + info["prob"].q @ z)(_pack(info["prob"], Xs, Us))
    assert abs(obj(X, U) - info["cert"].objective) <= 1e-6 * max(1.0, abs(info["cert"].objective))


def _pack(prob, X, U):
    z = np.zeros(prob.nz)
    z[prob.ix] = X
    z[prob.iu] = U
    return z


def test_qp_parity_per_particle_slew_through_handle_api(gpu):
    """Per-particle slew weights (expressible in the C ABI, c_interface.jl:44-46) and state boxes."""
    rng = np.random.default_rng(11)
    for (M, N, n, m, Nc) in [(4, 6, 4, 2, 1), (3, 5, 4, 2, 0)]:
        args, kw = rand_problem(rng, M, N, n, m, True, 0.2)
        Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
        x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
        with DeviceSolver(M, N, n, m, Nc) as s:
            s.set_problem(x0, Q, R, Xr, Ur, None, None, kw["lu"], kw["uu"], kw["reg_x"], kw["reg_u"], kw["slew_reg"],
                          kw["slew_reg0"], kw["slew_um1"])
            s.set_warm_start(Xp, Up)
            s.set_linearization(f, fx, fu)
            s.qp_solve(keep_prev=True)
            X, U = s.solution()
        assert rel(X, Xo) < 1e-7 and rel(U, Uo) < 1e-7


def test_state_box_active_is_solved_by_admm(gpu):
    rng = np.random.default_rng(100)
    M, N, n, m, Nc = 3, 6, 4, 2, 1
    args, kw = rand_problem(rng, M, N, n, m, True, 0.5, per_particle_slew=False)
    kw["slew_reg0"] = np.full(M, 0.7)
    Xo, _, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
    xb = 0.85 * np.abs(Xo).max()
    kw.update(lx=-xb * np.ones((M, N, n)), ux=xb * np.ones((M, N, n)))
    Xo, Uo, info = lqp.lqp_solve(*args, Nc=Nc, **kw)
    assert np.sum(np.abs(Xo) >= xb - 1e-9) >= 1
    X, U = _aff(args, kw, Nc)
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5
    assert np.abs(X).max() <= xb * (1 + 1e-6)


def test_sentinel_sweep_double_integrator(gpu):
    """Every present/absent (NaN) combination the reference sweeps through c_lqp_solve
    (c_precompile.jl:52-100), compared with the oracle."""
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
                    Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **k)
                    X, U = _aff(args, k, Nc)
                    assert rel(X, Xo) < 1e-7 and rel(U, Uo) < 1e-7, (ub, xb, s0, s)


def test_cone_entry_point_single_particle(gpu):
    """c_lcone_solve with hard constraints and M = 1 has the QP's minimiser."""
    args, kw, Nc = double_integrator(u_limit=0.4)
    Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    j = static_backend.py2jl
    X, U = static_backend.lcone_solve(Nc, j(x0), j(f), j(fx, 2), j(fu, 2), j(Xp), j(Up), j(Q, 2), j(R, 2), j(Xr), j(Ur),
                                      j(kw["lx"]), j(kw["ux"]), j(kw["lu"]), j(kw["uu"]), 1.0, 0.1, kw["slew_reg"],
                                      kw["slew_reg0"], j(kw["slew_um1"]), smooth_alpha=np.nan, solver="ecos")
    assert rel(X, Xo) < 1e-7 and rel(U, Uo) < 1e-7


def test_failure_convention_nan_in_gives_none_out(gpu):
    d, p = config_c1()
    d["Q"] = d["Q"].copy()
    d["Q"][0, 3, 0, 0] = np.nan
    assert pmpc_b200.solve(DubinsCar(p[0]), **d) == (None, None, None)


# ---------------------------------------------------------------------------------------- K1
def test_linearisation_kernel(gpu):
    """Device Dubins f, fx, fu: 1e-12 relative vs the oracle everywhere, and vs the reference's own
    autograd values where those are well conditioned (|u| >= 0.05)."""
    g = np.load(os.path.join(GOLD, "dubins_lin.npz"))
    for k in ("generic", "zero_u", "small_u"):
        X, U, P = g[k + "_X"], g[k + "_U"], g[k + "_p"]
        car = DubinsCar(P[:, 0, :])
        f, fx, fu = car(X, U)
        fo, fxo, fuo = dubins.car_f_fx_fu(X, U, P)
        assert rel(f, fo) < 1e-12 and rel(fx, fxo) < 1e-12 and rel(fu, fuo) < 1e-12
        if k == "generic":
            assert rel(f, g[k + "_f"]) < 1e-12 and rel(fx, g[k + "_fx"]) < 1e-12 and rel(fu, g[k + "_fu"]) < 1e-12


# ---------------------------------------------------------------------------------------- SCP
@pytest.mark.parametrize("name,kw", [
    ("scp_c1", dict(max_it=10, res_tol=-1.0, solver_settings=dict(solver="osqp"))),
    ("scp_consensus", dict(max_it=40, res_tol=1e-7, solver_settings=dict(solver="osqp", Nc=1))),
    ("scp_slew", dict(max_it=60, res_tol=1e-7, slew_rate=0.5, solver_settings=dict(solver="osqp", Nc=2, slew_reg=0.3))),
])
def test_host_loop_replays_reference_golden(gpu, name, kw):
    """pmpc_b200.solve with a Python callable that replays the linearisations the REFERENCE loop
    consumed must land on the reference loop's X, U and history (golden from make_golden.py)."""
    d = np.load(os.path.join(GOLD, name + ".npz"))
    it = iter(range(len(d["lin_f"])))

    def replay(X, U):
        k = next(it)
        return d["lin_f"][k], d["lin_fx"][k], d["lin_fu"][k]

    extra = {}
    if name == "scp_slew":
        extra = dict(u0_slew=0.2 * np.ones((3, 2)), x_l=d["x_l"], x_u=d["x_u"])
    X, U, data = pmpc_b200.solve(replay, d["Q"], d["R"], d["x0"], d["X_ref"], d["U_ref"], d["X_prev"], d["U_prev"],
                                 u_l=d["u_l"], u_u=d["u_u"], reg_x=1.0, reg_u=1e-2, **kw, **extra)
    hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
    assert len(hist) == len(d["hist"])
    assert rel(X, d["X"]) < 1e-5 and rel(U, d["U"]) < 1e-5
    assert np.abs(hist[:, 0] - d["hist"][:, 0]).max() <= 1e-6 * np.abs(d["hist"][:, 0]).max()


def _oracle_scp(d, p):
    fn = dubins.make_f_fx_fu_fn(p)
    kw = {k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "reg_x", "reg_u", "max_it", "res_tol")}
    return oscp.scp_solve(fn, d["Q"], d["R"], d["x0"], solver_settings=d["solver_settings"], **kw)


def test_device_scp_loop_c1_matches_oracle(gpu):
    """BASELINE config 1 fully on the device (linearise + QP + bookkeeping in one launch)."""
    d, p = config_c1()
    Xo, Uo, do = _oracle_scp(d, p)
    X, U, data = pmpc_b200.solve(DubinsCar(p[0]), **{k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in d.items()})
    assert X.shape == (21, 4) and U.shape == (20, 2) and len(data["hist"]) == 10
    assert rel(X, Xo[0]) < 1e-5 and rel(U, Uo[0]) < 1e-5
    ho = np.array([[h["obj"], h["resid"]] for h in do["hist"]])
    h = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
    assert np.abs(h[:, 0] - ho[:, 0]).max() <= 1e-6 * np.abs(ho[:, 0]).max()
    assert np.abs(h[:, 1] - ho[:, 1]).max() <= 1e-6 * np.abs(ho[:, 1]).max()


def test_device_scp_consensus_matches_oracle_and_host_loop(gpu):
    d, p = config_c2(M=8, N=15)
    d.update(res_tol=1e-7, max_it=60)
    Xo, Uo, do = _oracle_scp(d, p)
    X, U, data = pmpc_b200.solve(DubinsCar(p), **d)
    assert len(data["hist"]) == len(do["hist"])
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5
    assert abs(data["hist"][-1]["obj"] - do["hist"][-1]["obj"]) <= 1e-6 * abs(do["hist"][-1]["obj"])
    assert np.all(U[:, :1] == U[0:1, :1])  # consensus on the first Nc controls
    Xh, Uh, dh = pmpc_b200.solve(DubinsCar(p), **dict(d, solver_settings=dict(d["solver_settings"], host_loop=True)))
    assert rel(Xh, X) < 1e-9 and rel(Uh, U) < 1e-9 and len(dh["hist"]) == len(data["hist"])


def test_full_size_c2_properties(gpu):
    """BASELINE config 2 at full size (M=128, N=50): properties that do not need the oracle."""
    d, p = config_c2()
    car = DubinsCar(p)
    X, U, data = pmpc_b200.solve(car, **d)
    dev = data["device"]
    assert dev["status"] == 0 and data["hist"][-1]["resid"] < 1e-5
    assert np.all(U[:, 0] == U[0, 0])                                   # consensus
    assert U.max() <= 1 + 1e-9 and U.min() >= -1 - 1e-9                 # box
    assert max(sd["kkt"] for sd in data["solver_data"]) < 1e-8          # every QP certified on device
    # fixed point: the converged trajectory satisfies the true nonlinear dynamics to ~res_tol
    f, _, _ = car(X[:, :-1], U)
    assert np.abs(f - X[:, 1:]).max() < 1e-4
    # idempotence: restarting from the solution converges immediately to the same point
    X2, U2, data2 = pmpc_b200.solve(car, **dict(d, X_prev=X[:, 1:], U_prev=U))
    assert len(data2["hist"]) <= 2 and rel(X2, X) < 1e-4


def test_batched_equals_individual(gpu):
    """BASELINE config 3 (reduced B): B problems in one launch == B separate solves."""
    B = 6
    d, p = config_c3(B=B)
    probs = [dict({k: (v[b] if isinstance(v, np.ndarray) and v.shape[0] == B else v) for k, v in d.items()},
                  f_fx_fu_fn=DubinsCar(p[b])) for b in range(B)]
    res = pmpc_b200.solve_problems(probs)
    for b in (0, 3, 5):
        X, U, data = pmpc_b200.solve(**probs[b])
        assert rel(res[b][0], X) < 1e-12 and rel(res[b][1], U) < 1e-12
        assert len(res[b][2]["hist"]) == len(data["hist"])


def test_ltv_c4_reduced_matches_oracle(gpu):
    """BASELINE config 4 (reduced M, N): 12-state LTV, slew, control + state boxes, Nc=1."""
    d, A, Bm = config_c4(M=3, N=12)
    fn = LTV(A, Bm)
    X, U, data = pmpc_b200.solve(fn, **d)
    Xo, Uo, do = oscp.scp_solve(lambda X_, U_: fn(X_, U_), d["Q"], d["R"], d["x0"], X_ref=d["X_ref"], U_ref=d["U_ref"],
                                X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], x_l=d["x_l"],
                                x_u=d["x_u"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=d["slew_rate"],
                                max_it=d["max_it"], res_tol=d["res_tol"], solver_settings=d["solver_settings"])
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5


def test_interior_point_fallback_matches_oracle(gpu):
    """polish_iters = 1 makes every QP whose first active-set guess is wrong leave the primal-dual
    active-set iteration immediately: the interior-point fallback (+ certifying polish pass) has to
    deliver the same QP minimisers, i.e. the same SCP trajectory as the oracle."""
    d, p = config_c2(M=8, N=15)
    d.update(res_tol=1e-7, max_it=60)
    Xo, Uo, do = _oracle_scp(d, p)
    X, U, data = pmpc_b200.solve(DubinsCar(p), **dict(d, solver_settings=dict(d["solver_settings"], b200=dict(polish_iters=1))))
    assert data["device"]["status"] == 0 and data["device"]["admm_iters"] > 0  # fallback iterations are counted here
    assert len(data["hist"]) == len(do["hist"])
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5
    assert max(sd["kkt"] for sd in data["solver_data"]) < 1e-8
    # single-particle problems (thread-per-problem kernel) take the same route
    d1, p1 = config_c1()
    Xo1, Uo1, _ = _oracle_scp(d1, p1)
    k1 = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in d1.items()}
    X1, U1, data1 = pmpc_b200.solve(DubinsCar(p1[0]), **dict(k1, solver_settings=dict(d1["solver_settings"], b200=dict(polish_iters=1))))
    assert rel(X1, Xo1[0]) < 1e-5 and rel(U1, Uo1[0]) < 1e-5


def _gpu_count():
    from pmpc_b200 import _lib
    return _lib.device_count()


def test_particle_sharded_consensus_two_gpus(gpu, tmp_path):
    """SURVEY 8e case 2: ONE consensus problem, particles sharded over 2 GPUs, consensus terms exchanged
    inside the kernel over peer memory == the same problem on one GPU."""
    import json
    import subprocess
    import sys

    if _gpu_count() < 2:
        # one-GPU box: the same exchange between two handles of this process (their kernels are co-resident on the one
        # device; two PROCESSES would be time-sliced, which a spin-wait exchange cannot live with).  The IPC plumbing
        # itself needs the second GPU.
        from pmpc_b200 import parallel

        for (M, N, Nc, box) in [(64, 20, 1, 1.0), (64, 20, 2, 5.0)]:
            d, p = config_c2(M, N)
            d["u_l"], d["u_u"] = box * d["u_l"], box * d["u_u"]
            kw = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                      reg_x=d["reg_x"], reg_u=d["reg_u"], max_it=d["max_it"], res_tol=d["res_tol"], Nc=Nc)
            X, U, infos = parallel.solve_consensus_sharded_local(2, p, d["Q"], d["R"], d["x0"], **kw)
            kw1 = dict({k: v for k, v in kw.items() if k != "Nc"}, solver_settings=dict(d["solver_settings"], Nc=Nc))
            X1, U1, data1 = pmpc_b200.solve(DubinsCar(p), d["Q"], d["R"], d["x0"], **kw1)
            assert all(i["status"] == 0 and i["scp_iters"] == len(data1["hist"]) for i in infos)
            assert np.all(U[:, :Nc] == U[:1, :Nc]) and np.abs(X - X1).max() < 1e-8 and np.abs(U - U1).max() < 1e-8
        return
    out = tmp_path / "sharded.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(os.path.dirname(__file__), "sharded_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(out.read_text())
    for name, v in res.items():
        assert v["status"] == 0 and v["consensus"], (name, v)
        assert v["scp_iters"] == v["scp_iters_single"], (name, v)
        assert v["err_X"] < 1e-8 and v["err_U"] < 1e-8, (name, v)


# ---- user device dynamics (pmpc_register_dynamics: CUDA source compiled with NVRTC) ---------------------
EULER_CAR_SRC = r"""
// Euler-discretised unicycle: state (x, y, v, th), control (acc, turn), p = (a_scale, w_scale, T)
__device__ void euler_car(const double* x, const double* u, const double* p, double* f, double* fx, double* fu) {
  const double T = p[2], s = sin(x[3]), c = cos(x[3]);
  f[0] = x[0] + T * x[2] * c;  f[1] = x[1] + T * x[2] * s;  f[2] = x[2] + T * p[0] * u[0];  f[3] = x[3] + T * p[1] * u[1];
  fx[0 * 4 + 0] = 1.0; fx[0 * 4 + 2] = T * c; fx[0 * 4 + 3] = -T * x[2] * s;
  fx[1 * 4 + 1] = 1.0; fx[1 * 4 + 2] = T * s; fx[1 * 4 + 3] = T * x[2] * c;
  fx[2 * 4 + 2] = 1.0; fx[3 * 4 + 3] = 1.0;
  fu[2 * 2 + 0] = T * p[0]; fu[3 * 2 + 1] = T * p[1];
}
"""
PENDULUM_SRC = r"""
// damped pendulum, semi-implicit Euler: state (th, w), control torque, no parameters
__device__ void pend(const double* x, const double* u, const double* p, double* f, double* fx, double* fu) {
  const double T = 0.1, w = x[1] + T * (-9.81 * sin(x[0]) - 0.2 * x[1] + 2.0 * u[0]);
  f[0] = x[0] + T * w;  f[1] = w;
  fx[1 * 2 + 0] = -T * 9.81 * cos(x[0]);  fx[1 * 2 + 1] = 1.0 - 0.2 * T;
  fx[0 * 2 + 0] = 1.0 + T * fx[1 * 2 + 0];  fx[0 * 2 + 1] = T * fx[1 * 2 + 1];
  fu[1] = 2.0 * T;  fu[0] = T * fu[1];
}
"""


def _euler_car_np(p):
    p = np.asarray(p, float)

    def fn(X, U):
        pp = p[:, None, :] if (p.ndim == 2 and X.ndim == 3) else p
        a, w, T = pp[..., 0], pp[..., 1], pp[..., 2]
        s, c = np.sin(X[..., 3]), np.cos(X[..., 3])
        f = np.stack([X[..., 0] + T * X[..., 2] * c, X[..., 1] + T * X[..., 2] * s, X[..., 2] + T * a * U[..., 0],
                      X[..., 3] + T * w * U[..., 1]], -1)
        fx = np.zeros(X.shape + (4,)); fu = np.zeros(X.shape + (2,))
        for k in range(4):
            fx[..., k, k] = 1.0
        fx[..., 0, 2], fx[..., 0, 3] = T * c, -T * X[..., 2] * s
        fx[..., 1, 2], fx[..., 1, 3] = T * s, T * X[..., 2] * c
        fu[..., 2, 0], fu[..., 3, 1] = T * a, T * w
        return f, fx, fu

    return fn


def _pend_np(X, U):
    T = 0.1
    w = X[..., 1] + T * (-9.81 * np.sin(X[..., 0]) - 0.2 * X[..., 1] + 2.0 * U[..., 0])
    f = np.stack([X[..., 0] + T * w, w], -1)
    fx = np.zeros(X.shape + (2,)); fu = np.zeros(X.shape + (1,))
    fx[..., 1, 0], fx[..., 1, 1] = -T * 9.81 * np.cos(X[..., 0]), 1.0 - 0.2 * T
    fx[..., 0, 0], fx[..., 0, 1] = 1.0 + T * fx[..., 1, 0], T * fx[..., 1, 1]
    fu[..., 1, 0] = 2.0 * T
    fu[..., 0, 0] = T * fu[..., 1, 0]
    return f, fx, fu


def _euler_problem(M, N, seed, Nc=1):
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(0.8, 1.2, M), rng.uniform(0.8, 1.2, M), np.full(M, 0.2)], -1)
    d = dict(Q=np.tile(np.eye(4), (M, N, 1, 1)), R=np.tile(0.05 * np.eye(2), (M, N, 1, 1)),
             x0=np.tile(np.array([1.0, -0.5, 0.6, 0.3]) + 0.1 * rng.standard_normal(4), (M, 1)),
             X_ref=np.zeros((M, N, 4)), U_ref=np.zeros((M, N, 2)), X_prev=np.zeros((M, N, 4)), U_prev=np.zeros((M, N, 2)),
             u_l=-np.ones((M, N, 2)), u_u=np.ones((M, N, 2)), reg_x=1.0, reg_u=1e-2, max_it=80, res_tol=1e-7,
             solver_settings=dict(solver="osqp", Nc=Nc))
    return d, p


def test_user_dynamics_linearisation_matches_numpy(gpu):
    """The NVRTC-compiled user kernel (K1) against the same formulas in NumPy: 1e-12 relative (north_star)."""
    rng = np.random.default_rng(5)
    M, N = 5, 9
    p = np.stack([rng.uniform(0.8, 1.2, M), rng.uniform(0.8, 1.2, M), np.full(M, 0.2)], -1)
    X, U = rng.standard_normal((M, N, 4)), rng.standard_normal((M, N, 2))
    f, fx, fu = pmpc_b200.UserDynamics(EULER_CAR_SRC, "euler_car", 4, 2, p)(X, U)
    fo, fxo, fuo = _euler_car_np(p)(X, U)
    assert rel(f, fo) < 1e-12 and rel(fx, fxo) < 1e-12 and rel(fu, fuo) < 1e-12
    Xp, Up = rng.standard_normal((3, 7, 2)), rng.standard_normal((3, 7, 1))
    f, fx, fu = pmpc_b200.UserDynamics(PENDULUM_SRC, "pend", 2, 1)(Xp, Up)
    fo, fxo, fuo = _pend_np(Xp, Up)
    assert rel(f, fo) < 1e-12 and rel(fx, fxo) < 1e-12 and rel(fu, fuo) < 1e-12


def test_user_dynamics_scp_matches_oracle(gpu):
    """Whole SCP solve with registered device dynamics (consensus, box) == oracle with the NumPy callable, and ==
    the reference-style host loop that calls the same compiled kernel as an ordinary f_fx_fu_fn."""
    d, p = _euler_problem(M=6, N=12, seed=11)
    dyn = pmpc_b200.UserDynamics(EULER_CAR_SRC, "euler_car", 4, 2, p)
    X, U, data = pmpc_b200.solve(dyn, **d)
    kw = {k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "reg_x", "reg_u", "max_it", "res_tol")}
    Xo, Uo, do = oscp.scp_solve(_euler_car_np(p), d["Q"], d["R"], d["x0"], solver_settings=d["solver_settings"], **kw)
    assert len(data["hist"]) == len(do["hist"]) and data["hist"][-1]["resid"] < 1e-7
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5
    assert abs(data["hist"][-1]["obj"] - do["hist"][-1]["obj"]) <= 1e-6 * abs(do["hist"][-1]["obj"])
    assert np.all(U[:, :1] == U[0:1, :1])
    Xh, Uh, dh = pmpc_b200.solve(dyn, **dict(d, solver_settings=dict(d["solver_settings"], host_loop=True)))
    assert rel(Xh, X) < 1e-9 and rel(Uh, U) < 1e-9 and len(dh["hist"]) == len(data["hist"])
    # single particle, other shape (2 states, 1 control, no parameters)
    N = 25
    dp = dict(Q=np.tile(np.diag([4.0, 0.5]), (N, 1, 1)), R=np.tile(0.1 * np.eye(1), (N, 1, 1)), x0=np.array([2.5, 0.0]),
              X_prev=np.zeros((N, 2)), U_prev=np.zeros((N, 1)), u_l=-np.ones((N, 1)), u_u=np.ones((N, 1)), max_it=100,
              res_tol=1e-8, reg_x=1.0, reg_u=1e-1, solver_settings=dict(solver="osqp"))
    X, U, data = pmpc_b200.solve(pmpc_b200.UserDynamics(PENDULUM_SRC, "pend", 2, 1), **dp)
    Xo, Uo, do = oscp.scp_solve(_pend_np, **dp)
    assert X.shape == (N + 1, 2) and len(data["hist"]) == len(do["hist"])
    assert rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5


def test_user_dynamics_batched_equals_individual_and_errors(gpu):
    """B problems with registered dynamics in one handle (host-driven loop, converged problems frozen) == B separate
    solves, although they stop after different numbers of iterations; a source that does not compile raises with the
    NVRTC log."""
    B = 5
    probs = []
    for b in range(B):
        d, p = _euler_problem(M=4, N=10, seed=20 + b)
        d["res_tol"] = 1e-6
        d["x0"] = d["x0"] * (0.2 + 0.4 * b)   # different distances from the origin -> different iteration counts
        probs.append(dict(d, f_fx_fu_fn=pmpc_b200.UserDynamics(EULER_CAR_SRC, "euler_car", 4, 2, p)))
    res = pmpc_b200.solve_problems(probs)
    its = [len(r[2]["hist"]) for r in res]
    assert len(set(its)) > 1, its
    for b in range(B):
        X, U, data = pmpc_b200.solve(**probs[b])
        assert rel(res[b][0], X) < 1e-12 and rel(res[b][1], U) < 1e-12 and len(data["hist"]) == its[b]
    with pytest.raises(pmpc_b200.PmpcError, match="failed to compile"):
        pmpc_b200.UserDynamics("__device__ void broken(const double* x) { undeclared = 1; }", "broken", 2, 1)(
            np.zeros((1, 3, 2)), np.zeros((1, 3, 1)))


def test_pipelined_batch_equals_single_launch(gpu):
    """solve_batch(chunks=3): sub-batches on their own handles / streams with uploads overlapping the solve kernels
    must return bit-identical X, U and counters (problems are independent), also with page-locked reuse."""
    B = 7
    d, p = config_c3(B=B, M=5, N=12)
    args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=0.0, Nc=1)
    Q, R = np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"])
    X1, U1, info1, _ = pmpc_b200.solve_batch(p, Q, R, d["x0"], **args)
    pipe = None
    for rep in range(2):  # second call reuses the pipeline (handles + page-locked outputs)
        X3, U3, info3, pipe = pmpc_b200.solve_batch(p, Q, R, d["x0"], chunks=3, solver=pipe, reuse_buffers=True, **args)
        assert np.array_equal(X1, X3) and np.array_equal(U1, U3)
        assert all(np.array_equal(info1[k], info3[k]) for k in info1)


def test_accelerated_outer_loop_against_reference_golden(gpu):
    """`accelerated_scp_solve` (momentum over SCP iterates, reference pmpc/accelerated.py:15-116) with the device
    Dubins model against the trajectory the REFERENCE's own function produced here (make_golden_accelerated.py)."""
    g = np.load(os.path.join(GOLD, "scp_accelerated.npz"))
    kw = {k: g[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u")}
    X, U, data = pmpc_b200.accelerated_scp_solve(DubinsCar(g["p"][:, 0]), g["Q"], g["R"], g["x0"], verbose=False, max_it=60,
                                                 res_tol=1e-7, reg_x=1.0, reg_u=1e-2, solver_settings=dict(solver="osqp", Nc=1), **kw)
    assert abs(len(data["hist"]) - len(g["hist"])) <= 1 and data["hist"][-1]["resid"] < 1e-7
    assert rel(X, g["X"]) < 1e-5 and rel(U, g["U"]) < 1e-5
    assert abs(data["hist"][-1]["obj"] - g["hist"][-1, 0]) <= 1e-6 * abs(g["hist"][-1, 0])
    k = min(len(data["hist"]), len(g["hist"])) - 2  # same contraction history away from the noisy first iterates (H3)
    assert abs(data["hist"][k]["resid"] - g["hist"][k, 1]) <= 0.05 * g["hist"][k, 1] + 1e-9


def test_tune_scp_picks_the_smallest_residual(gpu):
    d, p = config_c2(M=4, N=10)
    d.update(max_it=15, res_tol=1e-9)
    grid = np.array([1e-2, 1e0, 1e2])
    reg_x, reg_u = pmpc_b200.tune_scp(DubinsCar(p), d["Q"], d["R"], d["x0"], reg_list=grid, reg_ratio=1e-2,
                                      **{k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "max_it", "res_tol",
                                                           "solver_settings")})
    res = []
    for r in grid:
        _, _, data = pmpc_b200.solve(DubinsCar(p), **dict(d, reg_x=float(r), reg_u=float(r) * 1e-2))
        res.append(data["hist"][-1]["resid"])
    assert reg_x == grid[int(np.argmin(res))] and reg_u == pytest.approx(1e-2 * reg_x)


def test_thread_per_problem_batch_with_ragged_tail(gpu):
    """B = 35 single-particle problems (one thread per problem: a full warp + a 3-lane tail whose idle lanes exit at
    once) == the oracle on each problem; the kernel must not depend on what the idle lanes would have done."""
    B, N = 35, 8
    d, p = config_c3(B=B, M=1, N=N)
    probs = [dict({k: (v[b, 0] if isinstance(v, np.ndarray) and v.shape[:2] == (B, 1) else v) for k, v in d.items()},
                  f_fx_fu_fn=DubinsCar(p[b, 0])) for b in range(B)]
    for q in probs:
        q["solver_settings"] = dict(solver="osqp")
        q["res_tol"] = 1e-7
    res = pmpc_b200.solve_problems(probs)
    for b in (0, 17, 31, 32, 34):
        q = probs[b]
        kw = {k: q[k][None] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u")}
        Xo, Uo, do = oscp.scp_solve(dubins.make_f_fx_fu_fn(p[b, 0]), q["Q"][None], q["R"][None], q["x0"][None], reg_x=q["reg_x"],
                                    reg_u=q["reg_u"], max_it=q["max_it"], res_tol=q["res_tol"], solver_settings=q["solver_settings"], **kw)
        assert res[b][0].shape == (N + 1, 4) and len(res[b][2]["hist"]) == len(do["hist"])
        assert rel(res[b][0], Xo[0]) < 1e-5 and rel(res[b][1], Uo[0]) < 1e-5


@pytest.mark.parametrize("M,N,Nc", [(1024, 3, 1), (2, 1, 1), (70, 6, 6)])
def test_extreme_shapes_properties(gpu, M, N, Nc):
    """Largest particle count per GPU (M = 1024, kind-2 CTAs), a one-stage horizon, full consensus (Nc = N) on a
    ragged particle count: consensus equality, box, exact linearised dynamics at the fixed point, device KKT certificate."""
    d, p = config_c2(M=M, N=N)
    d.update(res_tol=1e-8, max_it=80, solver_settings=dict(solver="osqp", Nc=Nc))
    car = DubinsCar(p)
    X, U, data = pmpc_b200.solve(car, **d)
    assert data["device"]["status"] == 0 and data["hist"][-1]["resid"] < 1e-8
    assert np.all(U[:, :Nc] == U[:1, :Nc]) and np.abs(U).max() <= 1 + 1e-12
    assert max(sd["kkt"] for sd in data["solver_data"]) < 1e-8
    f, _, _ = car(X[:, :-1], U)
    assert np.abs(f - X[:, 1:]).max() < 1e-6


def test_random_shape_sweep_qp_parity(gpu):
    """Seeded sweep over problem shapes the fixed CASES do not hit (odd particle counts around the warp size, every
    compiled kernel shape via zero-padding, Nc from 0 to N, slew on / off, box on / off): one convex sub-problem through
    the C ABI against the exact oracle, 1e-8."""
    rng = np.random.default_rng(2026)
    worst = 0.0
    for trial in range(24):
        n = int(rng.integers(1, 7)); m = int(rng.integers(1, min(n, 3) + 1))
        M = int(rng.choice([1, 2, 3, 7, 31, 32, 33, 65])); N = int(rng.integers(1, 9))
        Nc = int(rng.choice([-1, 0, 1, 2, N])); Nc = min(Nc, N)
        slew = bool(rng.integers(0, 2)); box = [None, 0.15, 0.4][int(rng.integers(0, 3))]
        args, kw = rand_problem(rng, M, N, n, m, slew, box, None, per_particle_slew=False)
        if slew:
            kw["slew_reg0"] = np.full(M, float(rng.uniform(0.3, 1.2)))  # the Python path broadcasts one scalar
        Xo, Uo, _ = lqp.lqp_solve(*args, Nc=Nc, **kw)
        X, U = _aff(args, kw, Nc)
        err = max(np.abs(X - Xo).max(), np.abs(U - Uo).max())
        worst = max(worst, err)
        assert err < 1e-8, (trial, (M, N, n, m, Nc, slew, box), err)
    assert worst < 1e-8


def test_full_size_c3_batch_properties(gpu):
    """BASELINE config 3 at its full per-problem size (N=30, M=16, Nc=1), 1024 of the 16384 problems in one launch
    (two particle-problems' worth of idle lanes per warp: M=16 in 32-lane tiles): status, stop rule, consensus, box,
    nonlinear-dynamics fixed point; three problems re-solved individually must agree bit for bit."""
    B = 1024
    d, p = config_c3(B=B)
    args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=0.0, Nc=1)
    X, U, info, ds = pmpc_b200.solve_batch(p, np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"]), d["x0"], **args)
    assert (info["status"] == 0).all() and (info["scp_iters"] < d["max_it"]).all()
    assert np.all(U[:, :, :1] == U[:, :1, :1]) and np.abs(U).max() <= 1 + 1e-12 and np.isfinite(X).all()
    for b in (0, 511, 1023):
        car = DubinsCar(p[b])
        f, _, _ = car(X[b, :, :-1], U[b])
        assert np.abs(f - X[b, :, 1:]).max() < 1e-4
        X1, U1, data = pmpc_b200.solve(car, d["Q"][b], d["R"][b], d["x0"][b], X_ref=d["X_ref"][b], U_ref=d["U_ref"][b],
                                       X_prev=d["X_prev"][b], U_prev=d["U_prev"][b], u_l=d["u_l"][b], u_u=d["u_u"][b],
                                       max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"],
                                       solver_settings=d["solver_settings"])
        assert np.array_equal(X1, X[b]) and np.array_equal(U1, U[b]) and len(data["hist"]) == info["scp_iters"][b]


def test_full_size_c4_properties(gpu):
    """BASELINE config 4 at full size (n=12, m=4, N=100, M=64, slew, control and state boxes, Nc=1): the LTV system is
    exactly linear, so the solution must satisfy the dynamics to rounding; consensus, boxes, device KKT certificate."""
    d, A, Bm = config_c4()
    X, U, data = pmpc_b200.solve(LTV(A, Bm), **d)
    assert data["device"]["status"] == 0 and data["hist"][-1]["resid"] < 1e-6
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    assert np.abs(f - X[:, 1:]).max() < 1e-12
    assert np.all(U[:, :1] == U[:1, :1]) and np.abs(U).max() <= 1 + 1e-12 and np.abs(X).max() <= 20 + 1e-9
    assert max(sd["kkt"] for sd in data["solver_data"]) < 1e-8


def test_lti_device_loop_matches_oracle(gpu):
    """LTI system (one A, B for the whole horizon; north_star's "LTI/LTV") through the device loop, with consensus, slew and
    a control box, against the oracle's SCP loop with the same linear model as a plain callable."""
    rng = np.random.default_rng(8)
    M, N, n, m = 5, 14, 2, 1
    A, Bm = np.array([[1.0, 0.1], [0.0, 1.0]]), np.array([[0.005], [0.1]])
    dyn = pmpc_b200.LTI(A, Bm)
    d = dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(0.1 * np.eye(m), (M, N, 1, 1)), x0=np.tile(np.array([3.0, -1.0]), (M, 1)) + 0.2 * rng.standard_normal((M, n)),
             X_prev=np.zeros((M, N, n)), U_prev=np.zeros((M, N, m)), u_l=-0.5 * np.ones((M, N, m)), u_u=0.5 * np.ones((M, N, m)),
             reg_x=1.0, reg_u=1e-2, slew_rate=0.2, max_it=60, res_tol=1e-8, solver_settings=dict(solver="osqp", Nc=2))
    X, U, data = pmpc_b200.solve(dyn, **d)
    Xo, Uo, do = oscp.scp_solve(lambda X_, U_: dyn(X_, U_), **d)
    assert len(data["hist"]) == len(do["hist"]) and rel(X, Xo) < 1e-5 and rel(U, Uo) < 1e-5
    assert np.all(U[:, :2] == U[:1, :2])

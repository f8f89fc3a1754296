"""Development check of the warp-per-particle kernel (wpp.cuh) against the thread-per-particle kernels of the same
shapes (solver_settings b200 fused=0 selects the old path): small LTV problems, BASELINE config 4 at full size and
config 5 reduced / full.     usage: python scripts/dev_wpp_check.py [small] [c4] [c5r] [c5]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from pmpc_b200 import LTV
from fixtures import config_c4, config_c5, config_ltv  # noqa

which = sys.argv[1:] or ["small", "c4", "c5r"]


def run(d, fn, fused, reps=1):
    dd = dict(d, solver_settings=dict(d["solver_settings"], b200=dict(fused=fused)))
    out = None
    ts = []
    for _ in range(reps + 1):
        t = time.perf_counter(); out = pmpc_b200.solve(fn, **dd); ts.append(time.perf_counter() - t)
    return out, min(ts[1:]) if reps else ts[0]


def report(label, d, A, Bm, old=True, reps=1):
    fn = LTV(A, Bm)
    (X, U, data), t = run(d, fn, 1, reps)
    row = dict(case=label)
    if data is None:
        row.update(failed=True)
        print(json.dumps(row), flush=True)
        return
    dev = data["device"]
    Nc = d["solver_settings"]["Nc"]
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    row.update(wall_s=t, kernel_ms=dev["kernel_ms"], its=len(data["hist"]), status=dev["status"], passes=dev["polish_passes"],
               sweeps=dev["sweeps"], admm=dev["admm_iters"], resid=data["hist"][-1]["resid"], dyn=float(np.abs(f - X[:, 1:]).max()),
               consensus=bool(np.all(U[:, :Nc] == U[:1, :Nc])), umax=float(np.abs(U).max()))
    if old:
        (Xo, Uo, do), to = run(d, fn, 0, reps)
        row.update(old_kernel_ms=do["device"]["kernel_ms"], old_its=len(do["hist"]), dX=float(np.abs(X - Xo).max()),
                   dU=float(np.abs(U - Uo).max()), old_passes=do["device"]["polish_passes"])
    print(json.dumps(row), flush=True)


if "small" in which:
    for (n, m, M, N, Nc, box, xbox) in [(8, 4, 1, 6, 0, 0.3, False), (8, 4, 3, 6, 1, 0.3, False), (12, 4, 5, 9, 2, 0.3, False),
                                        (16, 8, 9, 7, 1, 0.3, False), (32, 8, 4, 8, 5, 0.3, False), (32, 8, 20, 12, 2, 1.0, False),
                                        (12, 4, 64, 20, 1, 0.5, True), (20, 6, 33, 10, 3, 0.3, False)]:
        d, A, Bm = config_ltv(7 + n + M, M, N, n, m, Nc, xbox=xbox)
        d["u_l"], d["u_u"] = box * d["u_l"], box * d["u_u"]
        d["x0"] = 3.0 * d["x0"]
        report("n%d m%d M%d N%d Nc%d" % (n, m, M, N, Nc), d, A, Bm)
    # no slew (the lean instantiation)
    d, A, Bm = config_ltv(99, 6, 10, 12, 4, 2, xbox=False)
    d["slew_rate"] = 0.0
    d["u_l"], d["u_u"] = 0.3 * d["u_l"], 0.3 * d["u_u"]
    report("n12 m4 M6 N10 Nc2 no slew", d, A, Bm)
if "c4" in which:
    d, A, Bm = config_c4()
    report("C4 full", d, A, Bm, reps=3)
if "c5r" in which:
    d, A, Bm = config_c5(M=128, N=200)
    report("C5 M=128 N=200", d, A, Bm, old=bool(int(os.environ.get("C5R_OLD", "1"))))
if "c5" in which:
    d, A, Bm = config_c5()
    report("C5 full", d, A, Bm, old=False, reps=2)
if "c4lin" in which:
    # BASELINE config 4 as specified: with n_lin = 8 random linear rows per (particle, stage) (host loop, rows as extra states)
    M4, N4 = int(os.environ.get("C4_M", 64)), int(os.environ.get("C4_N", 100))
    d, A, Bm = config_c4(M=M4, N=N4, n_lin=8)
    fn = LTV(A, Bm)
    if "C4_IT" in os.environ:
        d["max_it"] = int(os.environ["C4_IT"])
    t = time.perf_counter()
    X, U, data = pmpc_b200.solve(fn, **d)
    t = time.perf_counter() - t
    Gu, Gx, h = d["solver_settings"]["lin_cstrs"]
    val = np.einsum("ijka,ija->ijk", Gu, U) + np.einsum("ijkr,ijr->ijk", Gx, X[:, 1:])
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    print(json.dumps(dict(case="C4 M=%d N=%d with n_lin=8" % (M4, N4), wall_s=t, its=len(data["hist"]), resid=data["hist"][-1]["resid"],
                          obj=data["hist"][-1]["obj"], t_aff_solve_mean=float(np.mean(data["t_aff_solve"])),
                          row_violation=float((val - h).max()), active_rows=int(np.sum(val >= h - 1e-7)),
                          dyn=float(np.abs(f - X[:, 1:]).max()), consensus=bool(np.all(U[:, :1] == U[:1, :1])),
                          umax=float(np.abs(U).max()))), flush=True)

if "c4box" in which:
    # config 4 at full size with a state box that binds on the second half of the horizon (tests: test_active_state_box_at_c4_shape)
    d, A, Bm = config_c4(M=64, N=100)
    d["x0"] = 2.0 * d["x0"]
    fn = LTV(A, Bm)
    free = dict(d); free.pop("x_l"); free.pop("x_u")
    Xf, _, _ = pmpc_b200.solve(fn, **free)
    half = 50
    xb = 0.7 * np.abs(Xf[:, 1 + half:]).max()
    d["x_l"], d["x_u"] = -1e3 * np.ones_like(d["x_l"]), 1e3 * np.ones_like(d["x_u"])
    d["x_l"][:, half:], d["x_u"][:, half:] = -xb, xb
    for mu in ("30", "0"):
        os.environ["PMPC_AL_MU"] = mu
        for rep in range(2):
            t = time.perf_counter()
            X, U, data = pmpc_b200.solve(fn, **d)
            t = time.perf_counter() - t
        dev = data["device"]
        print(json.dumps(dict(case="C4 full size, active state box, PMPC_AL_MU=" + mu, wall_s=t, kernel_ms=dev.get("kernel_ms"), its=len(data["hist"]),
                              active=int(np.sum(np.abs(X[:, 1 + half:]) >= xb - 1e-8)), xmax_over_box=float(np.abs(X[:, 1 + half:]).max() / xb),
                              polish=int(np.sum(dev.get("polish_passes", 0))), admm=int(np.sum(dev.get("admm_iters", 0))), status=int(np.max(dev["status"])))), flush=True)
    os.environ.pop("PMPC_AL_MU")

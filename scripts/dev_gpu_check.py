"""Developer check (GPU): random QPs and Dubins SCP through the handle API vs precomputed oracle answers
(scripts/dev_gen_cases.py, run on the CPU box)."""
import sys, os, time, pickle
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import functools
print = functools.partial(print, flush=True)
from oracle import dubins
from pmpc_b200 import DeviceSolver
cases = pickle.load(open(os.path.join(os.path.dirname(__file__), "dev_cases.pkl"), "rb"))
worst = 0.0
for cs in cases:
    M, N, n, m, Nc = cs["dims"]; args, kw, Xo, Uo = cs["args"], cs["kw"], cs["Xo"], cs["Uo"]
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    s = DeviceSolver(M, N, n, m, Nc)
    s.set_problem(x0, Q, R, Xr, Ur, None, None, kw.get("lu"), kw.get("uu"), kw["reg_x"], kw["reg_u"],
                  kw.get("slew_reg"), kw.get("slew_reg0"), kw.get("slew_um1"))
    s.set_warm_start(Xp, Up)
    s.set_linearization(f, fx, fu)
    s.qp_solve(keep_prev=True)
    X, U = s.solution()
    inf = s.info()
    eX, eU = np.abs(X - Xo).max(), np.abs(U - Uo).max()
    worst = max(worst, eX, eU)
    print((M, N, n, m, Nc), "errX %.2e errU %.2e" % (eX, eU), {k: int(v[0]) for k, v in inf.items()}, "ms %.3f" % s.last_kernel_ms())
    s.close()
print("worst", worst)

# Dubins SCP (golden consensus case) fully on device
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scp_consensus.npz"))
M, N = d["x0"].shape[0], d["Q"].shape[1]
s = DeviceSolver(M, N, 4, 2, 1)
s.set_problem(d["x0"], d["Q"], d["R"], d["X_ref"], d["U_ref"], None, None, d["u_l"], d["u_u"], 1.0, 1e-2)
s.set_warm_start(d["X_prev"], d["U_prev"])
s.set_dubins(d["p"].reshape(M, 3))
f, fx, fu = s.linearize()
X_ = np.concatenate([d["x0"][:, None], d["X_prev"][:, :-1]], 1)
fo, fxo, fuo = dubins.car_f_fx_fu(X_, d["U_prev"], d["p"].reshape(M, 1, 3))
print("lin err", np.abs(f - fo).max(), np.abs(fx - fxo).max(), np.abs(fu - fuo).max())
s.scp_solve(40, 1e-7)
X, U = s.solution()
inf = s.info(); h = s.hist(40)
print("scp consensus: its", inf["scp_iters"], "errX %.2e errU %.2e" % (np.abs(X - d["X"][:, 1:]).max(), np.abs(U - d["U"]).max()),
      {k: int(v[0]) for k, v in inf.items()}, "ms %.3f" % s.last_kernel_ms())
print("resid hist", h["resid"][:5, 0], d["hist"][:5, 1])

# C2 timing
rng = np.random.default_rng(1)
M, N, n, m = 128, 50, 4, 2
p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)
Q = np.tile(np.eye(n), (M, N, 1, 1)); R = np.tile(1e-2 * np.eye(m), (M, N, 1, 1)); x0 = np.ones((M, n))
for B in [1, 148, 592]:
    s = DeviceSolver(M, N, n, m, 1, B=B)
    t = time.time()
    s.set_problem(np.tile(x0, (B, 1, 1)), np.tile(Q, (B, 1, 1, 1, 1)), np.tile(R, (B, 1, 1, 1, 1)), None, None, None, None,
                  -np.ones((B, M, N, m)), np.ones((B, M, N, m)), 1.0, 1e-2)
    s.set_dubins(np.tile(p, (B, 1, 1)))
    t_set = time.time() - t
    for rep in range(3):
        s.set_warm_start(None, None)
        t = time.time(); s.scp_solve(100, 1e-5); t_solve = time.time() - t
    inf = s.info()
    print("C2 B=%d: set %.3fs solve wall %.3f ms kernel %.3f ms; scp its %d admm %d fact %d status %d -> %.1f solves/s" % (
        B, t_set, 1e3 * t_solve, s.last_kernel_ms(), inf["scp_iters"][0], inf["admm_iters"][0], inf["factorizations"][0],
        inf["status"].max(), B / (s.last_kernel_ms() * 1e-3)))
    if B == 1:
        X, U = s.solution()
        ref = np.load("/tmp/c2_proto.npz") if os.path.exists("/tmp/c2_proto.npz") else None
        if ref is not None:
            print("  vs proto errX %.2e errU %.2e" % (np.abs(X - ref["X"]).max(), np.abs(U - ref["U"]).max()))
    s.close()

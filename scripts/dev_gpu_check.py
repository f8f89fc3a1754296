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


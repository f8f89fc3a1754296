"""Multiplier method for active state bounds in the warp-per-particle kernel: linear rows at reduced config-4 shape."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from fixtures import config_c4
from pmpc_b200.dynamics import LTV
M, N, L = int(os.environ.get("AL_M", 3)), int(os.environ.get("AL_N", 10)), int(os.environ.get("AL_L", 2))
d, A, Bm = config_c4(M=M, N=N, n_lin=L)
Gu, Gx, _ = d["solver_settings"]["lin_cstrs"]
xt, x = np.empty((M, N, 12)), d["x0"].copy()
for j in range(N):
    x = np.einsum("mrt,mt->mr", A[:, j], x)
    xt[:, j] = x
h = 0.2 + np.abs(np.einsum("mjkr,mjr->mjk", Gx, xt))
fn = LTV(A, Bm)
ss = {k: v for k, v in d["solver_settings"].items() if k != "lin_cstrs"}
kw = {k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "x_l", "x_u", "reg_x", "reg_u", "slew_rate", "res_tol")}
out = {}
for mu in sys.argv[1:] or ["1e4", "0"]:
    os.environ["PMPC_AL_MU"] = mu
    t = time.perf_counter()
    X, U, data = pmpc_b200.solve(fn, d["Q"], d["R"], d["x0"], solver_settings=dict(ss, lin_cstrs=(Gu, Gx, h)), max_it=int(os.environ.get("AL_IT", 3)), **kw)
    t = time.perf_counter() - t
    out[mu] = (X, U)
    val = np.einsum("ijka,ija->ijk", Gu, U) + np.einsum("ijkr,ijr->ijk", Gx, X[:, 1:])
    sd = data["solver_data"][-1] if data.get("solver_data") else {}
    print(json.dumps(dict(mu=mu, wall=t, its=len(data["hist"]), resid=data["hist"][-1]["resid"], viol=float((val - h).max()),
                          active=int(np.sum(np.abs(val - h) < 1e-7)), solver_data={k: (v if np.isscalar(v) else str(v)) for k, v in dict(sd).items()} if isinstance(sd, dict) else str(sd))), flush=True)
ks = list(out)
if len(ks) == 2:
    print("difference between the two:", float(np.abs(out[ks[0]][0] - out[ks[1]][0]).max()), float(np.abs(out[ks[0]][1] - out[ks[1]][1]).max()))

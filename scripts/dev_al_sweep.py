"""Random sweep of the multiplier method (active state boxes in the warp-per-particle kernel) against the oracle."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from pmpc_b200 import LTV
from fixtures import config_ltv
from oracle import scp as oscp

def oracle(d, fn):
    return oscp.scp_solve(lambda X_, U_: fn(X_, U_), d["Q"], d["R"], d["x0"], X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"],
                          U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], x_l=d.get("x_l"), x_u=d.get("x_u"), reg_x=d["reg_x"], reg_u=d["reg_u"],
                          slew_rate=d["slew_rate"], max_it=d["max_it"], res_tol=d["res_tol"], solver_settings=d["solver_settings"])

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 16
bad = 0
for c in range(ncase):
    n, m = [(8, 4), (12, 4), (16, 8), (20, 6)][c % 4]
    M, N, Nc = int(rng.integers(1, 6)), int(rng.integers(6, 14)), int(rng.integers(0, 3))
    d, A, Bm = config_ltv(900 + c, M, N, n, m, Nc, xbox=False)
    d["x0"] = float(rng.uniform(1.5, 3.0)) * d["x0"]
    d["slew_rate"] = float(rng.choice([0.0, 1.0]))
    d["u_l"], d["u_u"] = 0.6 * d["u_l"], 0.6 * d["u_u"]
    d["max_it"] = 6
    fn = LTV(A, Bm)
    Xf, _, _ = oracle(d, fn)
    half = N // 2
    fac = float(rng.uniform(0.6, 0.95))
    xb = fac * np.abs(Xf[:, 1 + half:]).max()
    d["x_l"], d["x_u"] = -1e3 * np.ones((M, N, n)), 1e3 * np.ones((M, N, n))
    d["x_l"][:, half:], d["x_u"][:, half:] = -xb, xb
    try:
        Xo, Uo, do = oracle(d, fn)
    except Exception as e:
        print(json.dumps(dict(case=c, skipped="oracle: " + str(e)[:60]))); continue
    if Xo is None:
        print(json.dumps(dict(case=c, skipped="oracle returned None (infeasible)"))); continue
    X, U, data = pmpc_b200.solve(fn, **d)
    dev = data["device"]
    err = max(np.abs(X - Xo).max() / max(1.0, np.abs(Xo).max()), np.abs(U - Uo).max() / max(1.0, np.abs(Uo).max()))
    ok = err < 1e-5 and dev["status"] == 0
    bad += not ok
    print(json.dumps(dict(case=c, n=n, m=m, M=M, N=N, Nc=Nc, slew=d["slew_rate"], fac=round(fac, 2), active=int(np.sum(np.abs(Xo[:, 1 + half:]) >= xb - 1e-9)),
                          err=float(err), status=dev["status"], polish=dev["polish_passes"], admm=dev["admm_iters"], its=[len(data["hist"]), len(do["hist"])], ok=bool(ok))), flush=True)
print("failures:", bad)

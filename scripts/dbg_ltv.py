import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import pmpc_b200
from pmpc_b200 import LTV
from fixtures import config_c4
from oracle import scp as oscp
np.set_printoptions(linewidth=200, precision=4, suppress=True)
for xbox in (True, False):
  for slew in (1.0, 0.0):
    d, A, Bm = config_c4(M=3, N=12)
    d["slew_rate"] = slew
    if not xbox:
        d.pop("x_l"); d.pop("x_u")
    fn = LTV(A, Bm)
    X, U, data = pmpc_b200.solve(fn, **d)
    Xo, Uo, do = oscp.scp_solve(lambda X_, U_: fn(X_, U_), d["Q"], d["R"], d["x0"], X_ref=d["X_ref"], U_ref=d["U_ref"],
                                X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], x_l=d.get("x_l"),
                                x_u=d.get("x_u"), reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=d["slew_rate"],
                                max_it=d["max_it"], res_tol=d["res_tol"], solver_settings=d["solver_settings"])
    print("xbox", xbox, "slew", slew, "its", len(data["hist"]), len(do["hist"]), data["device"])
    print(" errX per (particle, stage):\n", np.abs(X - Xo).max(-1))
    print(" errU per (particle, stage):\n", np.abs(U - Uo).max(-1))
    print(" resid hist dev", [h["resid"] for h in data["hist"]][:6], "oracle", [h["resid"] for h in do["hist"]][:6])

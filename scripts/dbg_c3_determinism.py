"""Run the C3 batch twice (2048 problems) and compare bit for bit; also single launch vs pipelined."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from fixtures import config_c3
B = 2048
d, p = config_c3(B=B)
args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], max_it=d["max_it"],
            res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=0.0, Nc=1)
Q, R = np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"])
outs = []
for rep in range(2):
    X, U, info, ds = pmpc_b200.solve_batch(p, Q, R, d["x0"], **args)
    outs.append((X.copy(), U.copy(), info))
X3, U3, info3, _ = pmpc_b200.solve_batch(p, Q, R, d["x0"], chunks=8, **args)
print("run-to-run identical:", np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]),
      "| pipelined identical:", np.array_equal(outs[0][0], X3) and np.array_equal(outs[0][1], U3),
      "| status ok:", bool((info3["status"] == 0).all()), "| passes", int(outs[0][2]["polish_passes"].sum()), int(info3["polish_passes"].sum()))

"""Distribution of the per-problem work in BASELINE config 3 (sweeps = 2 x (active-set passes + interior-point iterations))."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from fixtures import config_c3
B = 16384
d, p = config_c3(B=B)
args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
            max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=0.0, Nc=1)
X, U, info, ds = pmpc_b200.solve_batch(p, np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"]), d["x0"], **args)
sw = info["sweeps"].astype(float)
ad = info["admm_iters"]
q = lambda a, qs: [float(np.percentile(a, x)) for x in qs]
print(json.dumps(dict(sweeps_mean=float(sw.mean()), sweeps_pct=q(sw, [50, 90, 99, 99.9, 100]), scp_its=q(info["scp_iters"], [50, 90, 99, 100]),
                      problems_with_ipm=int((ad > 0).sum()), ipm_iters_pct=q(ad[ad > 0], [50, 90, 100]) if (ad > 0).any() else None,
                      top=sorted(sw.tolist())[-8:], pair_max_over_mean=float(np.maximum(sw[0::2], sw[1::2]).mean() / sw.mean()))))

"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): consensus Dubins (fused lean path, ring,
shared_small), slew variant, thread-per-problem batch, LTV n=12, and registered user dynamics."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from pmpc_b200 import DubinsCar, LTV
from fixtures import config_c2, config_c3, config_c4
d, p = config_c2(M=40, N=8); d.update(max_it=4)
X, U, data = pmpc_b200.solve(DubinsCar(p), **d); print("consensus lean", data["device"])
X, U, data = pmpc_b200.solve(DubinsCar(p), **dict(d, slew_rate=0.3)); print("consensus slew", data["device"])
d3, p3 = config_c3(B=5, M=1, N=6)
probs = [dict({k: (v[b, 0] if isinstance(v, np.ndarray) and v.shape[:2] == (5, 1) else v) for k, v in d3.items()}, f_fx_fu_fn=DubinsCar(p3[b, 0]), max_it=3,
              solver_settings=dict(solver="osqp")) for b in range(5)]
print("thread-per-problem", [len(r[2]["hist"]) for r in pmpc_b200.solve_problems(probs)])
d4, A, Bm = config_c4(M=3, N=5); d4.update(max_it=2)
X, U, data = pmpc_b200.solve(LTV(A, Bm), **d4); print("ltv n=12", data["device"])
SRC = "__device__ void f(const double* x, const double* u, const double* p, double* f, double* fx, double* fu) { f[0] = x[0] + 0.1 * x[1]; f[1] = x[1] + 0.1 * (u[0] - sin(x[0])); fx[0] = 1; fx[1] = 0.1; fx[2] = -0.1 * cos(x[0]); fx[3] = 1; fu[1] = 0.1; }"
N = 6
X, U, data = pmpc_b200.solve(pmpc_b200.UserDynamics(SRC, "f", 2, 1), np.tile(np.eye(2), (N, 1, 1)), np.tile(0.1 * np.eye(1), (N, 1, 1)), np.array([1.0, 0.0]),
                             u_l=-np.ones((N, 1)), u_u=np.ones((N, 1)), max_it=3, solver_settings=dict(solver="osqp"))
print("user dynamics", data["device"])

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
# round 2: two problems per warp (forced: the batch-size rule would pick the one-per-warp kernel), incl. the tight box that
# needs the in-kernel interior point and an odd batch (one empty half), more pairs than resident warps is not needed for the queue path
os.environ["PMPC_HALF"] = "1"
for (B, M, N, box) in [(9, 16, 6, 1.0), (7, 5, 5, 0.2)]:
    dh, ph = config_c3(B=B, M=M, N=N)
    args = dict(X_ref=dh["X_ref"], U_ref=dh["U_ref"], X_prev=dh["X_prev"], U_prev=dh["U_prev"], u_l=box * dh["u_l"], u_u=box * dh["u_u"],
                max_it=4, res_tol=dh["res_tol"], reg_x=dh["reg_x"], reg_u=dh["reg_u"], slew_rate=0.0, Nc=1)
    X, U, info, _ = pmpc_b200.solve_batch(ph, np.ascontiguousarray(dh["Q"]), np.ascontiguousarray(dh["R"]), dh["x0"], **args)
    print("two per warp", B, M, N, int(info["status"].max()), int(info["polish_passes"].sum()), int(info["admm_iters"].sum()))
os.environ.pop("PMPC_HALF")
# warp-per-particle kernel with a binding state box (multiplier method) at config 4's shape
d4, A, Bm = config_c4(M=3, N=8); d4.update(max_it=3)
d4["x0"] = 2.0 * d4["x0"]
free = dict(d4); free.pop("x_l"); free.pop("x_u")
Xf, _, _ = pmpc_b200.solve(LTV(A, Bm), **free)
xb = 0.9 * np.abs(Xf[:, 5:]).max()  # binds on the second half of the horizon, feasible (tests: test_active_state_box_at_c4_shape)
d4["x_l"], d4["x_u"] = -1e3 * np.ones_like(d4["x_l"]), 1e3 * np.ones_like(d4["x_u"])
d4["x_l"][:, 4:], d4["x_u"][:, 4:] = -xb, xb
X, U, data = pmpc_b200.solve(LTV(A, Bm), **d4); print("ltv n=12, state box active", data["device"], float(np.abs(X[:, 5:]).max() / xb))

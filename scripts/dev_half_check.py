"""Two problems per warp (kind-3 kernels, M <= 16) against the one-problem-per-warp kernels (PMPC_HALF=0): results must
be bit-identical (same arithmetic, the reductions only skip the zeros of the mirror lanes); timing on BASELINE config 3."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from fixtures import config_c3


def run(B, M=16, N=30, no_half=False, reps=1, chunks=1, **over):
    os.environ["PMPC_HALF"] = "0" if no_half else "1"
    d, p = config_c3(B=B, M=M, N=N)
    d.update(over)
    args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=d.get("slew_rate", 0.0), Nc=1)
    Q, R = np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"])
    ds = None
    for _ in range(reps + 1):
        t = time.perf_counter()
        X, U, info, ds = pmpc_b200.solve_batch(p, Q, R, d["x0"], solver=ds, chunks=chunks, **args)
        t = time.perf_counter() - t
    return X.copy(), U.copy(), info, ds.last_kernel_ms(), t


for (B, M, N, over) in [] if "bigonly" in sys.argv else [(301, 16, 30, {}), (64, 9, 12, {}), (33, 2, 7, {}), (50, 16, 20, dict(slew_rate=0.3)),
                        (40, 16, 15, dict(u_l=-0.2 * np.ones((40, 16, 15, 2)), u_u=0.2 * np.ones((40, 16, 15, 2))))]:
    Xh, Uh, ih, msh, _ = run(B, M, N, False, **over)
    Xr, Ur, ir, msr, _ = run(B, M, N, True, **over)
    print(json.dumps(dict(case="B%d M%d N%d %s" % (B, M, N, sorted(over)), identical=bool(np.array_equal(Xh, Xr) and np.array_equal(Uh, Ur)),
                          its_equal=bool(np.array_equal(ih["scp_iters"], ir["scp_iters"])), status=[int(ih["status"].max()), int(ir["status"].max())],
                          passes=[int(ih["polish_passes"].sum()), int(ir["polish_passes"].sum())], admm=[int(ih["admm_iters"].sum()), int(ir["admm_iters"].sum())],
                          dmax=float(np.nanmax(np.abs(Xh - Xr))), n_differ=int(np.sum(np.abs(Xh - Xr).reshape(B, -1).max(1) > 0)),
                          ms=[msh, msr])), flush=True)
if "big" in sys.argv or "bigonly" in sys.argv:
    B = 16384
    _, _, ih, msh, th = run(B, reps=2)
    _, _, ir, msr, tr = run(B, no_half=True, reps=2)
    print(json.dumps(dict(case="C3 full", kernel_ms=[msh, msr], solves_per_s=[B / msh * 1e3, B / msr * 1e3], wall=[th, tr])), flush=True)

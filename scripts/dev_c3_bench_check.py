"""bench.py's C3 batch through the two-per-warp and the one-per-warp kernels: same iterates? who needs 100 iterations?"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
bench.set_workload("c3")
import pmpc_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
arrs, p, d0 = bench.make_batch(B, 0)
out = {}
for half in ("1", "0"):
    os.environ["PMPC_HALF"] = half
    X, U, info, ds = pmpc_b200.solve_batch(p, arrs["Q"], arrs["R"], arrs["x0"], arrs["X_ref"], arrs["U_ref"], arrs["X_prev"], arrs["U_prev"],
                                           None, None, arrs["u_l"], arrs["u_u"], max_it=100, res_tol=1e-5, reg_x=d0["reg_x"], reg_u=d0["reg_u"],
                                           slew_rate=0.0, Nc=1)
    hist = ds.hist(int(info["scp_iters"].max()))
    last = np.array([hist["resid"][info["scp_iters"][b] - 1, b] for b in range(B)])
    out[half] = (X.copy(), U.copy(), {k: np.array(v).copy() for k, v in info.items()}, last, ds.last_kernel_ms())
    bad = np.flatnonzero(last >= 1e-5)
    print(json.dumps(dict(half=half, ms=out[half][4], status_max=int(info["status"].max()), its_max=int(info["scp_iters"].max()),
                          not_converged=bad.tolist()[:20], n_bad=int(bad.size), last_bad=last[bad][:5].tolist(),
                          its_bad=info["scp_iters"][bad][:10].tolist())), flush=True)
    for b in bad[:2]:
        print("resid tail", b, hist["resid"][max(0, info["scp_iters"][b] - 8):info["scp_iters"][b], b].tolist())
(Xh, Uh, ih, lh, _), (Xr, Ur, ir, lr, _) = out["1"], out["0"]
dx = np.abs(Xh - Xr).reshape(B, -1).max(1)
print(json.dumps(dict(identical=bool(np.array_equal(Xh, Xr) and np.array_equal(Uh, Ur)), n_differ=int((dx > 0).sum()), dmax=float(dx.max()),
                      its_equal=bool(np.array_equal(ih["scp_iters"], ir["scp_iters"])), differ=np.flatnonzero(dx > 0)[:10].tolist())))

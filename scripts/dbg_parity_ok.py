"""Which of bench.py's parity_ok conditions fails, and is the device path run-to-run deterministic?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import bench, pmpc_b200
from pmpc_b200 import DeviceSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
M, N, NX, NU, NC = bench.M, bench.N, bench.NX, bench.NU, bench.NC
arrs, p, d0 = bench.make_batch(B, 0)
st = dict((k, int(v)) for k, v in (kv.split('=') for kv in os.environ.get('PMPC_SET', '').split(',') if kv))
ds = DeviceSolver(M, N, NX, NU, NC, B=B, settings=st or None)
ds.set_problem(arrs["x0"], arrs["Q"], arrs["R"], arrs["X_ref"], arrs["U_ref"], None, None, arrs["u_l"], arrs["u_u"], d0["reg_x"], d0["reg_u"], np.zeros(B * M))
ds.set_dubins(p)
outs = []
for rep in range(3):
    ds.set_warm_start(None, None)
    ds.scp_solve(100, 1e-5)
    info = ds.info(); X, U = ds.solution(); X, U = X.copy(), U.copy()
    hist = ds.hist(int(info["scp_iters"].max()))
    last = np.array([hist["resid"][info["scp_iters"][b] - 1, b] for b in range(B)])
    print("rep", rep, "status!=0:", np.nonzero(info["status"])[0][:10], "res>=tol:", np.nonzero(~(last < 1e-5))[0][:10], "finite", np.isfinite(X).all(),
          "umax", np.abs(U).max(), "consensus", np.all(U[:, :, 0] == U[:, :1, 0]), "passes", info["factorizations"].sum(), "admm", info["admm_iters"].sum(), flush=True)
    outs.append((X, U, info))
for rep in (1, 2):
    dX = np.abs(outs[rep][0] - outs[0][0]).reshape(B, -1).max(1)
    print("rep", rep, "vs 0: problems differing", np.nonzero(dX > 0)[0][:10], "max", dX.max(), "iters differ", np.nonzero(outs[rep][2]["scp_iters"] != outs[0][2]["scp_iters"])[0][:10])
Xe, Ue, infoe, ds = pmpc_b200.solve_batch(p, arrs["Q"], arrs["R"], arrs["x0"], arrs["X_ref"], arrs["U_ref"], arrs["X_prev"], arrs["U_prev"], None, None, arrs["u_l"], arrs["u_u"],
                                          max_it=100, res_tol=1e-5, reg_x=d0["reg_x"], reg_u=d0["reg_u"], slew_rate=0.0, Nc=NC, solver=ds, reuse_buffers=True)
dX = np.abs(Xe[:, :, 1:].reshape(B, -1) - outs[0][0].reshape(B, -1)).max(1)
print("e2e vs rep0: problems differing", np.nonzero(dX > 0)[0][:10], "max", dX.max(), "status", np.nonzero(infoe["status"])[0][:10])
bad = np.nonzero(dX > 0)[0]
for b in bad[:4]:
    print(b, {k: (int(outs[0][2][k][b]), int(infoe[k][b])) for k in infoe})

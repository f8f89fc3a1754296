"""Diagnose stragglers in the bench batch: which problems leave the fused active-set path."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import bench
from pmpc_b200 import DeviceSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
st = dict((k, int(v)) for k, v in (kv.split('=') for kv in os.environ.get('PMPC_SET', '').split(',') if kv))
arrs, p, d0 = bench.make_batch(B, 0)
ds = DeviceSolver(128, 50, 4, 2, 1, B=B, settings=st or None)
ds.set_problem(arrs["x0"], arrs["Q"], arrs["R"], arrs["X_ref"], arrs["U_ref"], None, None, arrs["u_l"], arrs["u_u"], d0["reg_x"], d0["reg_u"], np.zeros(B * 128))
ds.set_dubins(p)
for rep in range(2):
    ds.set_warm_start(None, None)
    ds.scp_solve(100, 1e-5)
    inf = ds.info()
    print("kernel %.1f ms" % ds.last_kernel_ms(), {k: (int(v.min()), float(v.mean()), int(v.max())) for k, v in inf.items()})
bad = np.nonzero(inf["admm_iters"] > 0)[0]
print("problems with ADMM:", bad[:20], "count", len(bad))
h = ds.hist(int(inf["scp_iters"].max()))
for b in bad[:3]:
    its = int(inf["scp_iters"][b])
    print("problem", b, {k: int(v[b]) for k, v in inf.items()})
    print(" admm per it:", h["admm_iters"][:its, b].astype(int).tolist())
    print(" resid:", ["%.1e" % v for v in h["resid"][:its, b]])
top = np.argsort(-inf["sweeps"])[:5]
print("most sweeps:", [(int(b), int(inf["sweeps"][b]), int(inf["scp_iters"][b])) for b in top])

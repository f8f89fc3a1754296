import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import pmpc_b200
from pmpc_b200 import DubinsCar, DeviceSolver
from fixtures import config_c1
d, p = config_c1()
kw = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in d.items()}
ds = DeviceSolver(1, 20, 4, 2, 0)
ds.set_problem(kw["x0"], kw["Q"], kw["R"], None, None, None, None, kw["u_l"], kw["u_u"], 1.0, 1e-2, np.zeros(1))
ds.set_dubins(p[0])
for (mi, rt) in [(10, 1e-5), (10, -1.0), (10, 0.0), (10, 1e-300), (10, -1e-5), (10, 1e-9), (3, 1e-5), (3, -1.0)]:
    ds.set_warm_start(None, None)
    t = time.perf_counter(); ds.scp_solve(mi, rt); t = time.perf_counter() - t
    inf = ds.info(); X, U = ds.solution()
    print("max_it %3d res_tol %8.1e: kernel %.3f ms wall %.3f ms its %d passes %d sweeps %d  |X| %.6f" % (mi, rt, ds.last_kernel_ms(), 1e3 * t, inf["scp_iters"][0], inf["polish_passes"][0], inf["sweeps"][0], np.abs(X).sum()), flush=True)

"""C2 (Dubins N=50 M=128 Nc=1) device solve, B concurrent problems, `reps` timed launches.  Used under ncu."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pmpc_b200 import DeviceSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(1)
M, N, n, m = 128, 50, 4, 2
p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)
Q = np.tile(np.eye(n), (B, M, N, 1, 1)); R = np.tile(1e-2 * np.eye(m), (B, M, N, 1, 1)); x0 = np.ones((B, M, n))
x0 = x0 + 0.05 * np.random.default_rng(7).standard_normal((B, 1, n))  # distinct problems
st = dict((k, int(v)) for k, v in (kv.split('=') for kv in os.environ.get('PMPC_SET', '').split(',') if kv))
s = DeviceSolver(M, N, n, m, 1, B=B, settings=st or None)
s.set_problem(x0, Q, R, None, None, None, None, -np.ones((B, M, N, m)), np.ones((B, M, N, m)), 1.0, 1e-2)
s.set_dubins(np.tile(p, (B, 1, 1)))
occ = s.occupancy(); print("occupancy: %d CTAs/SM x %d SMs" % occ, flush=True)
for rep in range(reps):
    s.set_warm_start(None, None)
    t = time.time(); s.scp_solve(100, 1e-5); wall = time.time() - t
    inf = s.info()
    print("C2 B=%d rep %d: wall %.3f ms kernel %.3f ms | scp its %d..%d admm %d fact %d polish %d status %d | %.1f solves/s" % (
        B, rep, 1e3 * wall, s.last_kernel_ms(), inf["scp_iters"].min(), inf["scp_iters"].max(), inf["admm_iters"].max(),
        inf["factorizations"].max(), inf["polish_passes"].max(), inf["status"].max(), B / (s.last_kernel_ms() * 1e-3)), flush=True)

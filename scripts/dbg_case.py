import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from fixtures import rand_problem
from oracle import lqp
from pmpc_b200 import DeviceSolver
np.set_printoptions(linewidth=200, precision=3, suppress=False)
CASES = [(4, 7, 3, 2, -1, True, 0.2, None), (3, 6, 4, 2, 2, True, 0.1, None), (40, 10, 4, 2, 1, True, 0.3, None), (5, 8, 4, 2, 3, False, 0.3, None)]
for case in CASES:
    M, N, n, m, Nc, slew, box, xbox = case
    rng = np.random.default_rng(100 + sum(x for x in case[:5]))
    args, kw = rand_problem(rng, M, N, n, m, slew, box, xbox, per_particle_slew=False)
    if slew:
        kw["slew_reg0"] = np.full(M, 0.7)
    Xo, Uo, info = lqp.lqp_solve(*args, Nc=Nc, **kw)
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    for st in (dict(ring=1), dict(ring=0), dict(fused=0)):
        with DeviceSolver(M, N, n, m, Nc, settings=st) as s:
            s.set_problem(x0, Q, R, Xr, Ur, None, None, kw.get("lu"), kw.get("uu"), kw["reg_x"], kw["reg_u"], kw.get("slew_reg"),
                          kw.get("slew_reg0"), kw.get("slew_um1"))
            s.set_warm_start(Xp, Up)
            s.set_linearization(f, fx, fu)
            s.qp_solve(keep_prev=True)
            X, U = s.solution()
            inf = {k: int(v[0]) for k, v in s.info().items()}
        print(case, st, "errX %.2e errU %.2e" % (np.abs(X - Xo).max(), np.abs(U - Uo).max()), inf)

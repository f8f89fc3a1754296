"""Developer check (GPU): random QPs and Dubins SCP through the handle API vs the CPU oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import lqp, dubins, scp as oscp

rng = np.random.default_rng(0)

def rand_problem(M, N, n, m, slew=True, box=0.1, xbox=None):
    spd = lambda k: (lambda a: a @ a.T / k + 0.5 * np.eye(k))(rng.standard_normal((k, k)))
    Q = np.stack([[spd(n) for _ in range(N)] for _ in range(M)])
    R = np.stack([[spd(m) for _ in range(N)] for _ in range(M)])
    fx = np.eye(n) + 0.2 * rng.standard_normal((M, N, n, n))
    fu = rng.standard_normal((M, N, n, m))
    f = rng.standard_normal((M, N, n))
    x0 = rng.standard_normal((M, n))
    X_prev, U_prev = rng.standard_normal((M, N, n)), 0.3 * rng.standard_normal((M, N, m))
    X_ref, U_ref = rng.standard_normal((M, N, n)), 0.3 * rng.standard_normal((M, N, m))
    kw = dict(reg_x=1.0, reg_u=0.1)
    if slew:
        kw.update(slew_reg=rng.uniform(0.5, 1.5, M), slew_reg0=rng.uniform(0.5, 1.5, M), slew_um1=rng.standard_normal((M, m)))
    if box:
        kw.update(lu=-box * np.ones((M, N, m)), uu=box * np.ones((M, N, m)))
    if xbox:
        kw.update(lx=-xbox * np.ones((M, N, n)), ux=xbox * np.ones((M, N, n)))
    return (x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref), kw

cases = [(1, 5, 2, 1, 0, True, 0.1), (3, 6, 4, 2, 2, True, 0.1), (4, 7, 3, 2, -1, True, 0.2), (2, 4, 3, 2, 1, False, 0.1),
         (2, 1, 2, 1, 1, True, 0.1), (5, 8, 4, 2, 1, True, None), (5, 8, 4, 2, 3, False, 0.3), (3, 6, 4, 2, 0, True, 0.1),
         (40, 10, 4, 2, 1, True, 0.3), (3, 5, 6, 3, 1, True, 0.2), (2, 6, 12, 4, 1, True, 0.2)]

import pickle, time
out = []
for (M, N, n, m, Nc, slew, box) in cases:
    args, kw = rand_problem(M, N, n, m, slew, box)
    t = time.time()
    Xo, Uo, info = lqp.lqp_solve(*args, Nc=Nc, **kw)
    print((M, N, n, m, Nc, slew, box), info["cert"].iterations, "%.2fs" % (time.time() - t), flush=True)
    out.append(dict(dims=(M, N, n, m, Nc), args=args, kw=kw, Xo=Xo, Uo=Uo))
pickle.dump(out, open(os.path.join(os.path.dirname(__file__), "dev_cases.pkl"), "wb"))

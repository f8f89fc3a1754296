"""Golden fixture for the momentum-accelerated outer loop, generated FROM THE REFERENCE ITSELF (build container only):
the unmodified ``pmpc.accelerated.accelerated_scp_solve`` (reference ``pmpc/accelerated.py:15-116``) is executed with the
reference's own Dubins model (``tests/dubins_car.py``) and ``pmpc.scp_mpc.aff_solve`` routed to the exact QP oracle
(the Julia/OSQP back end cannot run here)  ->  ``scp_accelerated.npz``.
    python tests/golden/make_golden_accelerated.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/tests")
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
import dubins_car as ref_car  # noqa: E402
import pmpc.scp_mpc as ref_scp  # noqa: E402
import pmpc.accelerated as ref_acc  # noqa: E402
from oracle import scp as oscp  # noqa: E402

ref_scp.aff_solve = lambda *a, **k: oscp.aff_solve(*a, **k)[:2] + ({},)
rng = np.random.default_rng(1)
M, N, n, m = 6, 12, 4, 2
p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)[:, None, :]


def f_fx_fu_fn(X, U):
    pp = np.broadcast_to(p, X.shape[:-1] + (3,))
    return ref_car.f_np(X, U, pp), ref_car.fx_np(X, U, pp), ref_car.fu_np(X, U, pp)


args = dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(1e-2 * np.eye(m), (M, N, 1, 1)), x0=np.ones((M, n)),
            X_ref=np.zeros((M, N, n)), U_ref=np.zeros((M, N, m)), X_prev=np.zeros((M, N, n)), U_prev=np.zeros((M, N, m)),
            u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)))
X, U, data = ref_acc.accelerated_scp_solve(f_fx_fu_fn, **args, verbose=False, max_it=60, res_tol=1e-7, reg_x=1.0, reg_u=1e-2,
                                           solver_settings=dict(solver="osqp", Nc=1))
hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
np.savez(os.path.join(HERE, "scp_accelerated.npz"), X=X, U=U, hist=hist, p=p, **args)
print("scp_accelerated.npz: its", len(hist), "final resid %.3e obj %.6e" % (hist[-1, 1], hist[-1, 0]))

"""Golden fixtures for the optional host-side hooks of the SCP loop, generated FROM THE REFERENCE ITSELF
(build container only: needs /root/reference):   python tests/golden/make_golden_hooks.py

The unmodified reference `pmpc.scp_mpc.scp_solve` is executed with its `aff_solve` routed to the exact QP oracle
(the Julia back end cannot run here), exactly as tests/golden/make_golden.py does, with

* `lin_cost_fn` (`_augment_cost`, pmpc/scp_mpc.py:171-185)                       -> scp_lin_cost.npz
* `filter_method="AA"` from iteration 5 (pmpc/scp_mpc.py:37-62, 376-387)         -> scp_filter_aa.npz
* `return_min_viol=True` with a residual that is not monotone (:419-421)         -> scp_min_viol.npz

The linearisations the reference loop consumed are stored so the tests can replay them through `pmpc_b200.solve`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/tests")

import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

import dubins_car as ref_car  # noqa: E402
import pmpc.scp_mpc as ref_scp  # noqa: E402

from oracle import scp as oscp  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from fixtures import hooks_lin_cost_fn as lin_cost_fn  # noqa: E402


def run(name, M, N, p, **kw):
    lin = []

    def f_fx_fu_fn(X, U):
        pp = np.broadcast_to(np.asarray(p, float), X.shape[:-1] + (3,))
        out = ref_car.f_np(X, U, pp), ref_car.fx_np(X, U, pp), ref_car.fu_np(X, U, pp)
        lin.append(out)
        return out

    def patched_aff_solve(*args, **kws):
        X, U, info = oscp.aff_solve(*args, **kws)
        return X, U, {}

    ref_scp.aff_solve = patched_aff_solve
    n, m = 4, 2
    args = dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(1e-2 * np.eye(m), (M, N, 1, 1)), x0=np.ones((M, n)),
                X_ref=np.zeros((M, N, n)), U_ref=np.zeros((M, N, m)), X_prev=np.zeros((M, N, n)),
                U_prev=np.zeros((M, N, m)), u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)), reg_x=1.0, reg_u=1e-2)
    args.update(kw)
    X, U, data = ref_scp.scp_solve(f_fx_fu_fn, **args)
    hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
    save = {k: v for k, v in args.items() if isinstance(v, np.ndarray)}
    save.update(X=X, U=U, hist=hist, p=np.asarray(p, float), lin_f=np.stack([l[0] for l in lin]),
                lin_fx=np.stack([l[1] for l in lin]), lin_fu=np.stack([l[2] for l in lin]))
    if "min_viol_sol" in data:
        save.update(mv_X=data["min_viol_sol"][0], mv_U=data["min_viol_sol"][1])
    np.savez_compressed(os.path.join(HERE, name), **save)
    print(name, "its", len(hist), "final resid %.3e obj %.6e" % (hist[-1, 1], hist[-1, 0]))


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    M = 4
    p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)[:, None, :]
    run("scp_lin_cost.npz", M, 10, p, max_it=40, res_tol=1e-7, lin_cost_fn=lin_cost_fn,
        solver_settings=dict(solver="osqp", Nc=1))
    run("scp_filter_aa.npz", M, 10, p, max_it=25, res_tol=1e-9, filter_method="AA", filter_window=4, filter_it0=5,
        solver_settings=dict(solver="osqp", Nc=1))
    # weak proximal terms from a far start: the second iterate moves further than the first (residuals 7.6, 8.6), so the
    # best-residual iterate is NOT the last one
    run("scp_min_viol.npz", M, 10, p, max_it=2, res_tol=-1.0, reg_x=1e-8, reg_u=1e-8, x0=3.0 * np.ones((M, 4)),
        return_min_viol=True, min_viol_it0=-1, solver_settings=dict(solver="osqp", Nc=1))

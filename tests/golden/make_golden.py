"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

What is executed from the reference (unmodified, imported from /root/reference):

* ``tests/dubins_car.py``: ``f_np / fx_np / fu_np`` (torch autograd Jacobians)
  -> ``dubins_lin.npz``  (pins oracle/dubins.py and the CUDA linearisation kernel)
* ``pmpc.scp_mpc.scp_solve`` — the real SCP outer loop — with its ``aff_solve``
  module attribute routed to ``oracle.scp.aff_solve`` (the exact QP oracle), because
  the Julia/OSQP back end cannot run here
  -> ``scp_c1.npz`` (BASELINE config 1), ``scp_consensus.npz`` (small C2-like case),
     ``scp_slew.npz`` (slew + state box), pinning oracle/scp.py's loop semantics.
* the deterministic double-integrator input fixture of
  ``PMPC.jl/src/c_precompile.jl:6-49`` / ``PMPC.jl/pmpcjl/test.py:171-226`` is
  re-created in ``tests/fixtures.py`` (inputs only; the reference stores no outputs).
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

import dubins_car as ref_car  # noqa: E402  (reference tests/dubins_car.py)
import pmpc.scp_mpc as ref_scp  # noqa: E402  (reference pmpc/scp_mpc.py)

from oracle import scp as oscp  # noqa: E402


def gen_dubins():
    rng = np.random.default_rng(1234)
    cases = {}
    M, N = 3, 7
    # (a) well-conditioned controls, per-particle parameters
    X = rng.standard_normal((M, N, 4))
    U = rng.uniform(0.05, 1.0, (M, N, 2)) * rng.choice([-1.0, 1.0], (M, N, 2))
    p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)[:, None, :]
    p = np.broadcast_to(p, (M, N, 3)).copy()
    cases["generic"] = (X, U, p)
    # (b) the default first iterate U = 0 and tiny controls (ill-conditioned)
    cases["zero_u"] = (rng.standard_normal((M, N, 4)), np.zeros((M, N, 2)), np.broadcast_to(np.array([1.0, 1.0, 0.3]), (M, N, 3)).copy())
    cases["small_u"] = (rng.standard_normal((M, N, 4)), 1e-4 * rng.standard_normal((M, N, 2)), np.broadcast_to(np.array([1.0, 1.0, 0.3]), (M, N, 3)).copy())
    out = {}
    for k, (X, U, p) in cases.items():
        out[k + "_X"], out[k + "_U"], out[k + "_p"] = X, U, p
        out[k + "_f"] = ref_car.f_np(X, U, p)
        out[k + "_fx"] = ref_car.fx_np(X, U, p)
        out[k + "_fu"] = ref_car.fu_np(X, U, p)
    np.savez(os.path.join(HERE, "dubins_lin.npz"), **out)
    print("dubins_lin.npz", {k: v.shape for k, v in out.items() if k.startswith("generic")})


def run_ref_scp(name, M, N, p, **kw):
    """Run the REFERENCE scp_solve with aff_solve routed to the exact QP oracle."""

    lin = []  # the linearisations the reference loop consumed, for bit-tight replay in tests

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
    Q = np.tile(np.eye(n), (M, N, 1, 1))
    R = np.tile(1e-2 * np.eye(m), (M, N, 1, 1))
    x0 = np.ones((M, n))
    args = dict(
        Q=Q, R=R, x0=x0, X_ref=np.zeros((M, N, n)), U_ref=np.zeros((M, N, m)),
        X_prev=np.zeros((M, N, n)), U_prev=np.zeros((M, N, m)),
        u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)), reg_x=1.0, reg_u=1e-2,
    )
    args.update(kw)
    X, U, data = ref_scp.scp_solve(f_fx_fu_fn, **args)
    hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
    save = {k: v for k, v in args.items() if isinstance(v, np.ndarray)}
    save.update(X=X, U=U, hist=hist, p=np.asarray(p, float))
    save.update(lin_f=np.stack([l[0] for l in lin]), lin_fx=np.stack([l[1] for l in lin]),
                lin_fu=np.stack([l[2] for l in lin]))
    np.savez(os.path.join(HERE, name), **save)
    print(name, "its", len(hist), "final resid %.3e obj %.6e" % (hist[-1, 1], hist[-1, 0]))


if __name__ == "__main__":
    gen_dubins()
    # BASELINE config 1: N=20 M=1, exactly 10 iterations (res_tol=-1)
    run_ref_scp("scp_c1.npz", 1, 20, np.array([1.0, 1.0, 0.3]), max_it=10, res_tol=-1.0,
                solver_settings=dict(solver="osqp"))
    # small C2-like consensus case: perturbed-dynamics particles, Nc=1
    rng = np.random.default_rng(1)
    M = 6
    p = np.stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)], -1)[:, None, :]
    run_ref_scp("scp_consensus.npz", M, 12, p, max_it=40, res_tol=1e-7,
                solver_settings=dict(solver="osqp", Nc=1))
    # slew-rate + u0_slew (needs solver_settings["slew_reg"]) + state box, Nc=2
    # (speed lower bound v >= -0.05 becomes active; looser position bounds keep every
    # intermediate linearised QP feasible)
    run_ref_scp("scp_slew.npz", 3, 10, p[:3], max_it=60, res_tol=1e-7, slew_rate=0.5,
                u0_slew=0.2 * np.ones((3, 2)),
                x_l=np.tile([-5.0, -5.0, -0.05, -5.0], (3, 10, 1)), x_u=5.0 * np.ones((3, 10, 4)),
                solver_settings=dict(solver="osqp", Nc=2, slew_reg=0.3))

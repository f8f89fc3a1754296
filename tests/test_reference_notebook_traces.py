"""Pins against numbers the reference's REAL solver stack produced.

The reference repository stores no expected outputs in its tests, but three of its notebooks keep the iteration tables
that `pmpc.solve` printed while running on PMPC.jl (Julia, ECOS / JuMP) for fully specified, deterministic problems
(tests/golden/make_notebook_traces.py transcribes them into tests/golden/notebook_traces.json; `obj` and `resid` with the 4
significant digits of the reference's `%8.3e` table, `pmpc/scp_mpc.py:402-417`).  Reproducing a table row needs the
linearisation, the assembled sub-problem (costs, proximal weights, slew penalty, dynamics rows, consensus layout, bounds),
its exact minimiser and the SCP bookkeeping to agree with the reference, iteration after iteration.

* `examples/gpu_solver.ipynb` cell 18 (one particle, hard control box): 50 rows.  The hard-box sub-problem is the one
  the CUDA path solves: the exact CPU oracle AND `pmpc_b200.solve` on the GPU are compared with the table.
* `tests/root_testing.ipynb` cell 4 (one particle, slew penalty 1e2, `smooth_alpha = 1e-1`), cell 11 (M = 20 particles,
  consensus over the first 5 controls, slew penalty, `smooth_alpha = 1`) and `tests/demo_cost_jax.ipynb` cell 3 (N = 30,
  R = 0, X_ref = 0.4, `smooth_alpha = 1e3`): those runs smoothed the box rows with the cone
  route's log barrier, so they are compared with `oracle.barrier`, the barrier variant on top of the SAME matrices
  `oracle.lqp.build_lqp` assembles for the QP — they pin everything but the treatment of the box rows.

Tolerances are the print precision (half a unit of the 4th digit: <= 5e-4 relative) wherever the reference's own solver
noise is below it, and stated row ranges where it is not (late iterations of the first table, where `resid` is a
difference of iterates at the level of ECOS's tolerance).
"""
import json
import os

import numpy as np
import pytest

from oracle import barrier, dubins
from oracle import scp as oscp

HERE = os.path.dirname(os.path.abspath(__file__))
TRACES = json.load(open(os.path.join(HERE, "golden", "notebook_traces.json")))
N, n, m = 20, 4, 2


def rel(ours, printed):
    return np.abs(np.asarray(ours) / np.asarray(printed) - 1.0)


def table(name):
    rows = TRACES[name]["rows"]
    return np.array([r["obj"] for r in rows]), np.array([r["resid"] for r in rows])


def hist(data):
    return np.array([h["obj"] for h in data["hist"]]), np.array([h["resid"] for h in data["hist"]])


def base(M=None):
    """Cells 3 / 6 / 10-11 of the notebooks: N = 20, Q = I, R = 1e-2 I, x0 = 1, references and warm start 0, |u| <= 1."""
    d = dict(Q=np.tile(np.eye(n), (N, 1, 1)), R=np.tile(1e-2 * np.eye(m), (N, 1, 1)), x0=np.ones(n), X_ref=np.zeros((N, n)),
             U_ref=np.zeros((N, m)), X_prev=np.zeros((N, n)), U_prev=np.zeros((N, m)), u_l=-np.ones((N, m)), u_u=np.ones((N, m)))
    if M is not None:
        d = {k: np.tile(v, (M,) + (1,) * v.ndim) for k, v in d.items()}
    return d


def check_gpu_solver_table(obj, resid):
    """The 50 rows of examples/gpu_solver.ipynb cell 18.  obj: print precision on every row.  resid: print precision over
    the first 12 rows; later the reference's iterates carry its cone solver's error (resid is a difference of iterates):
    within 2.5 % on every row but 30 / 31, where a control leaves its bound one iteration apart in the two runs."""
    ref_obj, ref_resid = table("gpu_solver_cpu_check")
    assert len(obj) == len(ref_obj) == 50
    assert rel(obj, ref_obj).max() <= 3.5e-4, rel(obj, ref_obj).max()
    r = rel(resid, ref_resid)
    assert r[:12].max() <= 5e-4, r[:12]
    late = np.delete(r, [29, 30])
    assert late.max() <= 2.5e-2, late.max()
    assert r[[29, 30]].max() <= 0.15


GPU_SOLVER = dict(reg_x=3.0, reg_u=1.0, max_it=50, res_tol=1e-5)  # opts of cell 11; the car of cell 4 has eps = 1e-3


def test_exact_oracle_reproduces_the_reference_table_of_gpu_solver_notebook():
    """Hard control box, one particle: the exact QP oracle inside the restated SCP loop against the table PMPC.jl printed."""
    fn = dubins.make_f_fx_fu_fn(np.array([1.0, 1.0, 0.3]), eps=1e-3)
    d = base()
    X, U, data = oscp.scp_solve(fn, d.pop("Q"), d.pop("R"), d.pop("x0"), **d, **GPU_SOLVER)
    check_gpu_solver_table(*hist(data))
    assert np.abs(U).max() <= 1.0 + 1e-9 and np.isclose(np.abs(U).max(), 1.0)  # the box is active in this problem


def test_barrier_oracle_reproduces_the_reference_table_with_slew_and_smoothing():
    """tests/root_testing.ipynb cell 4: T = 1 car, slew penalty 1e2, log-barrier box (alpha = 0.1), 20 iterations down to
    resid 1.022e-05 — every printed obj and resid at print precision."""
    ref_obj, ref_resid = table("root_testing_single")
    fn = dubins.make_f_fx_fu_fn(np.array([1.0, 1.0, 1.0]), eps=1e-6)
    d = base()
    X, U, data = oscp.scp_solve(fn, d.pop("Q"), d.pop("R"), d.pop("x0"), **d, reg_x=1.0, reg_u=1.0, slew_rate=1e2, max_it=20,
                                res_tol=1e-12, qp=lambda *a, **k: barrier.barrier_lqp_solve(*a, smooth_alpha=1e-1, **k))
    obj, resid = hist(data)
    assert len(obj) == len(ref_obj) == 20
    assert rel(obj, ref_obj).max() <= 5e-4 and rel(resid, ref_resid).max() <= 5e-4, (rel(obj, ref_obj).max(), rel(resid, ref_resid).max())
    # what this table pins beyond the first one: with a hard box instead of the barrier the first row is already off by 12 %,
    # without the slew penalty by 16 %
    for change in (dict(), dict(slew_rate=0.0)):
        d = base()
        qp = "exact" if not change else (lambda *a, **k: barrier.barrier_lqp_solve(*a, smooth_alpha=1e-1, **k))
        _, _, other = oscp.scp_solve(fn, d.pop("Q"), d.pop("R"), d.pop("x0"), **d, **dict(dict(reg_x=1.0, reg_u=1.0, slew_rate=1e2,
                                     max_it=1, res_tol=1e-12), **change), qp=qp)
        assert rel(other["hist"][0]["obj"], ref_obj[0]) > 0.1


def test_barrier_oracle_follows_the_reference_table_with_zero_control_cost():
    """tests/demo_cost_jax.ipynb cell 3 (its `solve_cpu` table): N = 30, R = 0 (the proximal weight reg_u = 1e-2 is the
    only curvature in the controls), X_ref = 0.4, default weights, barrier alpha = 1e3.  A looser pin, by the reference's
    own doing: it linearises with autograd through the eps = 1e-6 formula at U_prev = 0, where that formula loses ~6
    digits, and iterations 2-3 inherit the noise.  Same 26 iterations; obj at print precision from the 4th
    row on (the converged cost 0.6955 included) and within 3 % before; resid within 5 % on every row."""
    ref_obj, ref_resid = table("demo_cost_cpu")
    fn = dubins.make_f_fx_fu_fn(np.array([1.0, 1.0, 0.3]), eps=1e-6)
    Z = lambda d: np.zeros((30, d))
    X, U, data = oscp.scp_solve(fn, np.tile(np.eye(n), (30, 1, 1)), np.zeros((30, m, m)), np.ones(n), 0.4 * np.ones((30, n)), Z(m), Z(n),
                                Z(m), u_l=-np.ones((30, m)), u_u=np.ones((30, m)), max_it=100,
                                qp=lambda *a, **k: barrier.barrier_lqp_solve(*a, smooth_alpha=1e3, **k))
    obj, resid = hist(data)
    assert len(obj) == len(ref_obj) == 26
    assert rel(obj[3:], ref_obj[3:]).max() <= 3e-4 and rel(obj[:3], ref_obj[:3]).max() <= 3e-2
    assert rel(resid, ref_resid).max() <= 5e-2


def test_barrier_oracle_reproduces_the_reference_consensus_table():
    """tests/root_testing.ipynb cell 11: 20 particles with their own (v_scale, w_scale, T), consensus over the first 5
    controls, slew penalty, log-barrier box (alpha = 1): the same 23 iterations, obj at print precision, resid within
    0.3 %.  The notebook's solver name ('jump') predates the cost anchor of `main.jl:219-227`: the plain sum of the particle
    costs (the QP route's objective) under the barrier lands inside the printed digits; with the anchor of the current
    source the objective differs in the 4th digit (asserted both ways)."""
    ref_obj, ref_resid = table("root_testing_consensus")
    M = 20
    P = np.linspace(0.7, 1.0, M)
    fn = dubins.make_f_fx_fu_fn(np.stack([P, P, P], -1), eps=1e-6)

    def run(anchor_eps, max_it=50):
        d = base(M)
        return hist(oscp.scp_solve(fn, d.pop("Q"), d.pop("R"), d.pop("x0"), **d, reg_x=1.0, reg_u=1.0, slew_rate=1e2, max_it=max_it,
                                   res_tol=1e-5, solver_settings=dict(Nc=5),
                                   qp=lambda *a, **k: barrier.barrier_lqp_solve(*a, smooth_alpha=1.0, anchor_eps=anchor_eps, **k))[2])

    obj, resid = run(0.0)
    assert len(obj) == len(ref_obj) == 23  # stops at the same iteration (resid < 1e-5)
    assert rel(obj, ref_obj).max() <= 3e-4, rel(obj, ref_obj).max()
    assert rel(resid, ref_resid).max() <= 3e-3, rel(resid, ref_resid).max()
    obj_a, resid_a = run(barrier.COST_ANCHOR_EPS, max_it=6)
    assert rel(obj_a, ref_obj[:6]).max() <= 6e-4 and rel(resid_a, ref_resid[:6]).max() <= 3e-3


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_table_of_gpu_solver_notebook(gpu):
    """The product against the reference's own printed output: `pmpc_b200.solve` with the device dynamics (whole SCP loop
    in one kernel launch) and with a Python f_fx_fu_fn (reference-style host loop, `c_lqp_solve` per iteration) on the
    problem of examples/gpu_solver.ipynb cell 18 — same assertions as for the oracle."""
    import pmpc_b200

    for fn, extra in ((pmpc_b200.DubinsCar(p=np.array([1.0, 1.0, 0.3]), eps=1e-3), {}),
                      (dubins.make_f_fx_fu_fn(np.array([1.0, 1.0, 0.3]), eps=1e-3), dict(solver_settings=dict(solver="osqp")))):
        d = base()
        X, U, data = pmpc_b200.solve(fn, d.pop("Q"), d.pop("R"), d.pop("x0"), **d, **GPU_SOLVER, **extra)
        check_gpu_solver_table(*hist(data))
        assert X.shape == (N + 1, n) and np.isclose(np.abs(U).max(), 1.0)

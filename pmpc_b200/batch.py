"""Batched front-end: many independent problems in one device launch (BASELINE config 3).

Counterpart of the reference's ``solve_problems`` (serial, ``scp_mpc.py:504-511``) and of the
ZeroMQ task farm ``pmpc.remote.solve_problems`` (``remote.py:391-452``): instead of farming one
problem per CPU worker, B problems of identical shape become B CTAs of the same kernel.
"""
from __future__ import annotations

import gc
import time
from typing import Any, Dict, List, Optional

import numpy as np

from .dynamics import BuiltinDynamics, DubinsCar, LTV, UserDynamics
from .handle import DeviceSolver


_CACHED_BATCH_MAX_UNITS = 1 << 18  # particle-stages (B*M*N); beyond it a batch handle is released after the call


def _sig(p: Dict[str, Any]):
    dyn = p.get("f_fx_fu_fn")
    if not isinstance(dyn, (DubinsCar, LTV, UserDynamics)):
        return None
    if any(p.get(k) is not None and p.get(k) != "" for k in ("lin_cost_fn", "extra_cstrs_fns")) or p.get("filter_method", ""):
        return None
    ss = p.get("solver_settings") or {}
    # anything `solve` would route to its host loop (scp_mpc.scp_solve: device_ok) is solved by `solve`, one by one
    if (ss.get("host_loop", False) or p.get("debug", False) or p.get("return_min_viol", False)
            or float(p.get("time_limit", 1000.0)) < 1000.0):
        return None
    x0 = np.asarray(p["x0"])
    Q, R = np.asarray(p["Q"]), np.asarray(p["R"])
    present = tuple(p.get(k) is not None and np.size(p.get(k)) > 0 for k in ("x_l", "x_u", "u_l", "u_u"))
    user = (dyn.src, dyn.fn_name, dyn.params is None) if isinstance(dyn, UserDynamics) else None
    # an LTV model with and one without an affine term must not share a launch (the group's set_ltv call is decided once)
    affine = (getattr(dyn, "c", None) is None) if isinstance(dyn, LTV) else None
    b200 = tuple(sorted((ss.get("b200") or {}).items()))  # device solver settings select the kernel's behaviour
    cone = (str(ss.get("solver", "ecos")).lower(), "smooth_cstr" in ss, "smooth_alpha" in ss)
    return (type(dyn).__name__, user, affine, b200, cone, getattr(dyn, "eps", None), x0.ndim, Q.shape, R.shape, ss.get("Nc", -1),
            ss.get("slew_reg"), float(p.get("reg_x", 1.0)), float(p.get("reg_u", 1e-2)), p.get("slew_rate", 0.0),
            p.get("u0_slew") is None, present, int(p.get("max_it", 100)), float(p.get("res_tol", 1e-5)))


def solve_same_shape(problems: List[Dict[str, Any]], device: int = -1, solver: Optional[DeviceSolver] = None,
                     settings: Optional[dict] = None):
    """All problems must share :func:`_sig`.  Returns ``(list of (X, U, data), DeviceSolver)``."""
    p0 = problems[0]
    B = len(problems)
    single = np.asarray(p0["x0"]).ndim == 1
    Q0 = np.asarray(p0["Q"])
    M = 1 if single else Q0.shape[0]
    N, n = Q0.shape[-3], Q0.shape[-1]
    m = np.asarray(p0["R"]).shape[-1]
    ss = p0.get("solver_settings") or {}
    Nc = ss.get("Nc", -1)
    max_it, res_tol = int(p0.get("max_it", 100)), float(p0.get("res_tol", 1e-5))

    def stack(key, shape, default=None):
        vals = [p.get(key) for p in problems]
        if all(v is None or np.size(v) == 0 for v in vals):
            return default
        return np.stack([np.asarray(v, np.float64).reshape(shape) for v in vals])

    from .scp_mpc import _device_settings, _check_formulation

    _check_formulation(ss, M)
    settings = dict(settings or {}, **_device_settings(ss))  # solver_settings["b200"], as `solve` honours it
    ds = solver
    if ds is None or (ds.B, ds.M, ds.N, ds.n, ds.m, ds.Nc) != (B, M, N, n, m, Nc):
        ds = DeviceSolver(M, N, n, m, Nc, B=B, device=device)
    ds.set_settings(**settings)
    X_ref, U_ref = stack("X_ref", (M, N, n)), stack("U_ref", (M, N, m))
    slew_rate = p0.get("slew_rate", 0.0)
    ds.set_problem(stack("x0", (M, n)), stack("Q", (M, N, n, n)), stack("R", (M, N, m, m)), X_ref, U_ref,
                   stack("x_l", (M, N, n)), stack("x_u", (M, N, n)), stack("u_l", (M, N, m)), stack("u_u", (M, N, m)),
                   float(p0.get("reg_x", 1.0)), float(p0.get("reg_u", 1e-2)),
                   None if slew_rate is None else np.full(B * M, float(slew_rate)),
                   np.full(B * M, float(ss["slew_reg"])) if "slew_reg" in ss else None, stack("u0_slew", (M, m)))
    # warm start defaults follow scp_mpc.py:313-314 (X_prev = X_ref, U_prev = U_ref)
    Xp = stack("X_prev", (M, N, n), X_ref)
    Up = stack("U_prev", (M, N, m), U_ref)
    ds.set_warm_start(Xp, Up)
    dyn0 = p0["f_fx_fu_fn"]
    if isinstance(dyn0, DubinsCar):
        ds.set_dubins(np.stack([p["f_fx_fu_fn"].params_for(M) for p in problems]), dyn0.eps)
    elif isinstance(dyn0, UserDynamics):
        ds.register_dynamics(dyn0.src, dyn0.fn_name,
                             None if dyn0.params is None else np.stack([p["f_fx_fu_fn"].params_for(M) for p in problems]))
    else:
        bc = lambda a, tail: np.stack([np.broadcast_to(np.asarray(getattr(p["f_fx_fu_fn"], a)), (M, N) + tail) for p in problems])
        ds.set_ltv(bc("A", (n, n)), bc("B", (n, m)), None if dyn0.c is None else bc("c", (n,)))
    t = time.time()
    ds.scp_solve(max_it, res_tol)
    wall = time.time() - t
    X, U = ds.solution()
    X, U = X.reshape(B, M, N, n), U.reshape(B, M, N, m)
    info = ds.info()
    its_max = int(info["scp_iters"].max())
    h = ds.hist(max(its_max, 1))
    x0s = stack("x0", (M, n))
    # result records in the reference's format (one dict per SCP iteration, scp_mpc.py:409-417).  Built from Python lists:
    # indexing NumPy scalars one by one costs more host time per problem than the solve kernel does at C3's batch size
    obj, resid, admm = h["obj"].T.tolist(), h["resid"].T.tolist(), h["admm_iters"].T.tolist()
    status, iters = info["status"].tolist(), info["scp_iters"].tolist()
    dev_keys, dev_cols = list(info), [v.tolist() for v in info.values()]
    reg_x, reg_u = float(p0.get("reg_x", 1.0)), float(p0.get("reg_u", 1e-2))
    Xfull = np.concatenate([x0s[:, :, None, :], X], -2)
    out = []
    collecting = gc.isenabled()
    gc.disable()  # ~120 small acyclic dicts per problem: the cycle collector's passes over them were a third of this loop
    try:
        for b in range(B):
            if int(status[b]) & 1:
                out.append((None, None, None))
                continue
            its = int(iters[b])
            hist = [{"it": i, "elaps": wall * i / its, "obj": o, "resid": r, "reg_x": reg_x, "reg_u": reg_u}
                    for i, o, r in zip(range(1, its + 1), obj[b], resid[b])]
            data = dict(hist=hist, solver_data=[{"admm_iters": int(a)} for a in admm[b][:its]],
                        t_aff_solve=[wall / its] * its, device={k: int(c[b]) for k, c in zip(dev_keys, dev_cols)})
            out.append((Xfull[b, 0], U[b, 0], data) if single else (Xfull[b], U[b], data))
    finally:
        if collecting:
            gc.enable()
    return out, ds


def solve_batch(dyn_params, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None, x_l=None, x_u=None, u_l=None,
                u_u=None, max_it: int = 100, res_tol: float = 1e-5, reg_x: float = 1.0, reg_u: float = 1e-2,
                slew_rate: Optional[float] = 0.0, Nc: int = -1, dyn_eps: float = 1e-6, device: int = -1,
                solver=None, settings: Optional[dict] = None, reuse_buffers: bool = False, chunks: int = 1):
    """Array form of the batched solve: B independent Dubins-car particle problems given as stacked
    HOST arrays with a leading problem axis — ``dyn_params (B,M,3)``, ``Q (B,M,N,n,n)``, ``R (B,M,N,m,m)``,
    ``x0 (B,M,n)``, vectors ``(B,M,N,d)`` — solved with one CTA per problem.  Argument
    meaning and defaults are those of ``solve`` (reference ``scp_mpc.py:205-238``; warm-start default
    ``X_prev = X_ref``, ``:313-314``).  Returns ``X (B,M,N+1,n)`` including x0, ``U (B,M,N,m)``, ``info``
    (per-problem counters of ``pmpc_get_info``), and the solver state for reuse via ``solver=``.
    Inputs in page-locked memory (``pmpc_b200.pinned_empty`` / ``pinned_copy``) are copied at full PCIe
    speed; ``reuse_buffers=True`` returns X, U in page-locked buffers owned by the solver.

    ``chunks > 1`` pipelines the batch: it is cut into ``chunks`` contiguous sub-batches, each on its own handle
    (own CUDA stream); the host-to-device upload and packing of sub-batch k+1 and the device-to-host read of
    sub-batch k-1 overlap the solve kernel of sub-batch k, so PCIe time leaves the critical path except for the
    first upload and the last download.  Results are identical to ``chunks=1`` (problems are independent).

    Inputs that are the same for every particle need not be tiled (the reference's API makes the caller tile them,
    ``tests/simple.py:20-30``): ``Q, R, X_ref, U_ref, X_prev, U_prev`` and the bounds may be given as ONE block
    ``(N, ...)`` for the whole batch or one block per problem ``(B, N, ...)``, ``x0`` as ``(B, n)``; the device packing
    kernels expand them (``pmpc_set_broadcast``), which removes most of the host-to-device traffic of a batch."""
    Q = np.asarray(Q)
    B, M = np.asarray(dyn_params).shape[:2]
    N, n = Q.shape[-3], Q.shape[-1]
    m = np.asarray(R).shape[-1]
    if chunks > 1 and B >= chunks:
        return _solve_batch_pipelined(dyn_params, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it, res_tol,
                                      reg_x, reg_u, slew_rate, Nc, dyn_eps, device, solver, settings, reuse_buffers, chunks)
    ds = solver
    if not isinstance(ds, DeviceSolver) or (ds.B, ds.M, ds.N, ds.n, ds.m, ds.Nc) != (B, M, N, n, m, Nc):
        ds = DeviceSolver(M, N, n, m, Nc, B=B, device=device, settings=settings)
    _upload_and_launch(ds, dyn_params, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it, res_tol, reg_x, reg_u,
                       slew_rate, dyn_eps)
    X, U = ds.trajectory(reuse_buffers)  # stream-ordered after the solve; synchronises
    return X, U, ds.info(), ds


def _upload_and_launch(ds, dyn_params, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it, res_tol, reg_x, reg_u,
                       slew_rate, dyn_eps):
    B, M = ds.B, ds.M
    ds.set_problem(x0, Q, R, X_ref, U_ref, x_l, x_u, u_l, u_u, reg_x, reg_u,
                   None if slew_rate is None else np.full(B * M, float(slew_rate)))
    ds.set_warm_start(X_ref if X_prev is None else X_prev, U_ref if U_prev is None else U_prev)
    ds.set_dubins(dyn_params, dyn_eps)
    ds.scp_solve_async(max_it, res_tol)


class BatchPipeline:
    """Solver state of a pipelined ``solve_batch``: one handle per sub-batch and the page-locked output buffers."""

    def __init__(self, key, bounds, solvers):
        self.key, self.bounds, self.solvers, self.out = key, bounds, solvers, None

    def last_kernel_ms(self):
        return sum(s.last_kernel_ms() for s in self.solvers)


def _solve_batch_pipelined(dyn_params, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it, res_tol, reg_x, reg_u,
                           slew_rate, Nc, dyn_eps, device, solver, settings, reuse_buffers, chunks):
    from ._lib import pinned_empty

    B, M = np.asarray(dyn_params).shape[:2]
    N, n = Q.shape[-3], Q.shape[-1]
    m = np.asarray(R).shape[-1]
    key = (B, M, N, n, m, Nc, chunks, device)
    pipe = solver if isinstance(solver, BatchPipeline) and solver.key == key else None
    if pipe is None:
        bounds = [(B * k // chunks, B * (k + 1) // chunks) for k in range(chunks)]
        pipe = BatchPipeline(key, bounds, [DeviceSolver(M, N, n, m, Nc, B=hi - lo, device=device, settings=settings)
                                           for lo, hi in bounds])
    if reuse_buffers:
        if pipe.out is None:
            pipe.out = (pinned_empty((B, M, N + 1, n)), pinned_empty((B, M, N, m)))
        X, U = pipe.out
    else:
        X, U = np.empty((B, M, N + 1, n)), np.empty((B, M, N, m))
    def cut(a, lo, hi, full_ndim=4):
        """sub-batch [lo, hi) of an input: arrays without a problem axis (one block for the whole batch) pass through"""
        if a is None:
            return None
        a = np.asarray(a)
        return a if a.ndim <= full_ndim - 2 else a[lo:hi]

    # uploads run one after the other (each gets the full PCIe bandwidth); every solve is asynchronous on its own stream,
    # so sub-batch k+1 is uploaded and packed while sub-batch k is being solved
    for (lo, hi), ds in zip(pipe.bounds, pipe.solvers):
        _upload_and_launch(ds, np.asarray(dyn_params)[lo:hi], cut(Q, lo, hi, 5), cut(R, lo, hi, 5), np.asarray(x0)[lo:hi], cut(X_ref, lo, hi),
                           cut(U_ref, lo, hi), cut(X_prev, lo, hi), cut(U_prev, lo, hi), cut(x_l, lo, hi), cut(x_u, lo, hi),
                           cut(u_l, lo, hi), cut(u_u, lo, hi), max_it, res_tol, reg_x, reg_u, slew_rate, dyn_eps)
    infos = []
    for (lo, hi), ds in zip(pipe.bounds, pipe.solvers):
        ds.trajectory(out=(X[lo:hi], U[lo:hi]))  # waits for this sub-batch only; later ones keep computing
        infos.append(ds.info())
    info = {k: np.concatenate([i[k] for i in infos]) for k in infos[0]}
    return X, U, info, pipe


def solve_batched(problems: List[Dict[str, Any]], verbose: bool = False, device: int = -1, **kw):
    from .scp_mpc import solve

    results: List[Any] = [None] * len(problems)
    groups: Dict[Any, List[int]] = {}
    for i, p in enumerate(problems):
        groups.setdefault(_sig(p), []).append(i)
    for sig, idxs in groups.items():
        if sig is None:
            for i in idxs:
                results[i] = solve(**dict(problems[i], verbose=verbose))
            continue
        # a control loop re-solves the same batch every step (and a farm worker sees the same batch size again and
        # again): moderate batches keep their handle (device slab, streams, pinned buffers) in solve's cache
        group = [problems[i] for i in idxs]
        Q0 = np.asarray(group[0]["Q"])
        M = 1 if np.asarray(group[0]["x0"]).ndim == 1 else Q0.shape[0]
        N, n, m = Q0.shape[-3], Q0.shape[-1], np.asarray(group[0]["R"]).shape[-1]
        cached = None
        if 1 < len(group) * M * N <= _CACHED_BATCH_MAX_UNITS:
            from .scp_mpc import _cached_solver

            cached = _cached_solver(M, N, n, m, (group[0].get("solver_settings") or {}).get("Nc", -1), device, B=len(group))
        res, _ = solve_same_shape(group, device=device, solver=cached)
        for i, r in zip(idxs, res):
            results[i] = r
    return results

"""Public API: ``solve`` / ``scp_solve`` — the SCP MPC loop of reference ``pmpc/scp_mpc.py``.

Same signature (positional order ``f_fx_fu_fn, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u,
u_l, u_u``), same defaults, same return values ``(X, U, data)`` and array layouts
(``scp_mpc.py:205-442``).  Two execution paths, both ending in hand-written CUDA kernels:

* **device loop** — when ``f_fx_fu_fn`` is a :class:`pmpc_b200.dynamics.BuiltinDynamics`
  (Dubins car, LTV) and no host-side hook (filter, ``lin_cost_fn``, ``extra_cstrs_fns``) is
  requested, the whole loop (linearise → convex sub-problem → residual/objective → stop rule)
  runs inside one persistent kernel launch (``pmpc_scp_solve``);
* **host loop** — any Python ``f_fx_fu_fn``: the loop below follows the reference line by line
  and calls ``aff_solve`` → ``lqp_solve`` → ``pmpcjl.lqp_solve`` → ``c_lqp_solve`` (C ABI) once per
  iteration, exactly where the reference calls Julia (``scp_mpc.py:370``).
"""
from __future__ import annotations

import math
import threading
import time
from collections import OrderedDict
from copy import copy
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import lin_cstrs
from .dynamics import BuiltinDynamics, DubinsCar, LTV, UserDynamics
from .handle import DeviceSolver
from .static_backend import aff_solve
from .utils import TablePrinter, atleast_nd

_SOLVER_KEYS = ("rho0", "alpha", "eps_abs", "eps_rel", "kkt_tol", "adapt_tol", "max_admm_iter", "check_every", "polish",
                "polish_iters", "adapt_rho", "warm_start", "fused", "ring")


def print_fn(*args, **kwargs):
    return print(*args, **kwargs)


# -- fixed-point filters over the last iterates (reference scp_mpc.py:37-62) -----------------------
def _stack(Fs):
    return np.stack([np.reshape(f, -1) for f in Fs], -1)


def AA_method(Fs: List[np.ndarray]) -> np.ndarray:
    """Anderson-type mixing weights: least-squares combination of the last residuals."""
    F = _stack(Fs)
    dF = F[:, :-1] - F[:, -1:]
    k = dF.shape[-1]
    theta = np.linalg.solve(dF.T @ dF + 1e-10 * np.eye(k), -dF.T @ F[:, -1]).reshape(-1)
    return np.concatenate([theta, [1.0 - theta.sum()]])


def smooth_method(Fs: List[np.ndarray]) -> np.ndarray:
    k = len(Fs)
    return np.full(k, 1.0 / k)


def select_method(Fs: List[np.ndarray]) -> np.ndarray:
    """Weights inversely proportional to the squared residual norms (sum to one)."""
    F = _stack(Fs)
    k = F.shape[-1]
    Amat = np.zeros((k + 1, k + 1))
    Amat[:k, :k] = np.diag(np.linalg.norm(F, axis=0) ** 2)
    Amat[:k, k] = 1.0
    Amat[k, :k] = 1.0
    rhs = np.zeros(k + 1)
    rhs[k] = 1.0
    return np.linalg.solve(Amat, rhs)[:k]


FILTER_MAP = dict(smooth=smooth_method, select=select_method, AA=AA_method)


def _augment_cost(lin_cost_fn, X_prev, U_prev, Q, R, X_ref, U_ref, problems):
    """Fold a linearised extra cost into the references (reference scp_mpc.py:171-185)."""
    if lin_cost_fn is None:
        return X_ref, U_ref
    cx, cu = lin_cost_fn(X_prev, U_prev, problems)
    if cx is not None:
        X_ref = X_ref - np.linalg.solve(Q, np.asarray(cx)[..., None])[..., 0]
    if cu is not None:
        U_ref = U_ref - np.linalg.solve(R, np.asarray(cu)[..., None])[..., 0]
    return X_ref, U_ref


def _bmv(A, x):
    return (A @ x[..., None])[..., 0]


def _normalise(Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u):
    """Shape handling and defaults of scp_mpc.py:297-321."""
    x0 = np.array(x0, dtype=np.float64)
    Q, R = np.array(Q, dtype=np.float64), np.array(R, dtype=np.float64)
    single = x0.ndim == 1
    if single:
        assert R.ndim == 3 and Q.ndim == 3
        Q, R, x0 = atleast_nd(Q, 4), atleast_nd(R, 4), atleast_nd(x0, 2)
        X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u = [
            None if z is None else atleast_nd(np.asarray(z, dtype=np.float64), 3)
            for z in (X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u)]
    else:
        assert x0.ndim == 2 and R.ndim == 4 and Q.ndim == 4
    M, N, xdim = Q.shape[:3]
    udim = R.shape[-1]
    X_ref = np.zeros((M, N, xdim)) if X_ref is None else np.array(X_ref, dtype=np.float64).reshape(M, N, xdim)
    U_ref = np.zeros((M, N, udim)) if U_ref is None else np.array(U_ref, dtype=np.float64).reshape(M, N, udim)
    X_prev = X_ref if X_prev is None else np.array(X_prev, dtype=np.float64).reshape(M, N, xdim)
    U_prev = U_ref if U_prev is None else np.array(U_prev, dtype=np.float64).reshape(M, N, udim)
    x_l, x_u, u_l, u_u = [np.zeros((0, 0, 0)) if z is None else np.array(z, dtype=np.float64) for z in (x_l, x_u, u_l, u_u)]
    return single, (M, N, xdim, udim), Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u


def _device_settings(solver_settings):
    s = dict((solver_settings or {}).get("b200", {}))
    unknown = set(s) - set(_SOLVER_KEYS)
    if unknown:
        raise KeyError(f"unknown pmpc_b200 solver settings: {sorted(unknown)}")
    return s


def _check_formulation(solver_settings, M):
    """Same notice as the host path (``static_backend.aff_solve``): a cone / smoothed request (the reference's default
    ``solver="ecos"``, ``static_backend.py:242-253``) is served by the QP (sum-of-costs, hard box) formulation."""
    from . import static_backend as sb

    solver = str(solver_settings.get("solver", "ecos")).lower()
    smoothing = "smooth_cstr" in solver_settings or not math.isnan(float(solver_settings.get("smooth_alpha", math.nan)))
    if ((solver in sb._CONE_SOLVERS and M > 1) or smoothing) and "cone" not in sb._warned:
        sb._warned.add("cone")
        import warnings

        warnings.warn("pmpc_b200 solves the QP (sum-of-costs, hard box) formulation of lqp_solve; the requested "
                      "cone/smoothed route is served by it (see DESIGN.md, 'out of scope').")


# Handles are cached per (device, shape): a plain reference-style `solve(...)` call re-uses the resident slab, stream and
# events of the previous call of the same shape instead of paying cudaMalloc / cudaFree every time.
_HANDLE_CACHE: "OrderedDict[tuple, DeviceSolver]" = OrderedDict()
_HANDLE_CACHE_MAX = 8
_HANDLE_LOCK = threading.Lock()


def _cached_solver(M, N, xdim, udim, Nc, device=-1, B=1):
    key = (int(device), M, N, xdim, udim, Nc) + ((int(B),) if B != 1 else ())
    with _HANDLE_LOCK:
        ds = _HANDLE_CACHE.pop(key, None)
        if ds is None or not ds._h.value:
            ds = DeviceSolver(M, N, xdim, udim, Nc, B=B, device=device)
        _HANDLE_CACHE[key] = ds
        while len(_HANDLE_CACHE) > _HANDLE_CACHE_MAX:
            _, old = _HANDLE_CACHE.popitem(last=False)
            old.close()
    return ds


def clear_handle_cache():
    """Release the device memory of every cached handle."""
    with _HANDLE_LOCK:
        while _HANDLE_CACHE:
            _, ds = _HANDLE_CACHE.popitem()
            ds.close()


# -- device loop ------------------------------------------------------------------------------------
def _device_scp(dyn, dims, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it, res_tol, reg_x, reg_u,
                slew_rate, u0_slew, solver_settings, verbose, t0):
    M, N, xdim, udim = dims
    Nc = solver_settings.get("Nc", -1)
    _check_formulation(solver_settings, M)
    cache = solver_settings.get("_device_solver")  # explicit handle (kept for callers that manage their own)
    key = (M, N, xdim, udim, Nc)
    if isinstance(cache, DeviceSolver) and (cache.B, cache.M, cache.N, cache.n, cache.m, cache.Nc) == (1,) + key and cache._h.value:
        ds = cache
    else:
        ds = _cached_solver(M, N, xdim, udim, Nc)
    ds.set_settings(**_device_settings(solver_settings))  # defaults unless solver_settings["b200"] overrides them
    absent = lambda z: z is None or np.size(z) == 0
    slew_reg = None if slew_rate is None else np.full(M, float(slew_rate))
    slew_reg0 = np.full(M, float(solver_settings["slew_reg"])) if "slew_reg" in solver_settings else None
    ds.set_problem(x0, Q, R, X_ref, U_ref, None if absent(x_l) else x_l, None if absent(x_u) else x_u,
                   None if absent(u_l) else u_l, None if absent(u_u) else u_u, reg_x, reg_u, slew_reg, slew_reg0,
                   None if u0_slew is None else np.asarray(u0_slew, np.float64).reshape(M, udim))
    ds.set_warm_start(X_prev, U_prev)
    if isinstance(dyn, DubinsCar):
        ds.set_dubins(dyn.params_for(M), dyn.eps)
    elif isinstance(dyn, LTV):
        shp = (M, N)
        ds.set_ltv(np.broadcast_to(dyn.A, shp + dyn.A.shape[-2:]), np.broadcast_to(dyn.B, shp + dyn.B.shape[-2:]),
                   None if dyn.c is None else np.broadcast_to(dyn.c, shp + (xdim,)))
    elif isinstance(dyn, UserDynamics):
        ds.register_dynamics(dyn.src, dyn.fn_name, dyn.params_for(M))
    else:
        raise TypeError(f"no device implementation for {type(dyn).__name__}")
    t_k = time.time()
    ds.scp_solve(max_it, res_tol)
    t_k = time.time() - t_k
    X, U = ds.solution()
    info = ds.info()
    its = int(info["scp_iters"][0])
    h = ds.hist(max(its, 1))
    if int(info["status"][0]) & 1 or np.isnan(X).any() or np.isnan(U).any():
        if verbose:
            print_fn("Solver failed...")
        return None, None, None
    names = ["it", "elaps", "obj", "resid", "reg_x", "reg_u"]
    tp = TablePrinter(names, fmts=["%04d", "%8.3e", "%8.3e", "%8.3e", "%8.3e", "%8.3e"])
    elaps = time.time() - t0
    data = dict(solver_data=[], hist=[], t_aff_solve=[])
    if verbose:
        print_fn(tp.make_header())
    for it in range(its):
        vals = (it + 1, elaps * (it + 1) / its, float(h["obj"][it, 0]), float(h["resid"][it, 0]), reg_x, reg_u)
        if verbose:
            print_fn(tp.make_values(vals))
        data["hist"].append(dict(zip(names, vals)))
        data["solver_data"].append(dict(admm_iters=int(h["admm_iters"][it, 0]), kkt=float(h["kkt"][it, 0])))
        data["t_aff_solve"].append(t_k / its)
    if verbose:
        print_fn(tp.make_footer())
    data["device"] = dict(kernel_ms=ds.last_kernel_ms(), **{k: int(v[0]) for k, v in info.items()})
    data["device_solver"] = ds
    X = np.concatenate([x0[:, None, :], X.reshape(M, N, xdim)], -2)
    return X, U.reshape(M, N, udim), data


# -- SCP MPC ----------------------------------------------------------------------------------------
def scp_solve(
    f_fx_fu_fn: Callable,
    Q: np.ndarray,
    R: np.ndarray,
    x0: np.ndarray,
    X_ref: Optional[np.ndarray] = None,
    U_ref: Optional[np.ndarray] = None,
    X_prev: Optional[np.ndarray] = None,
    U_prev: Optional[np.ndarray] = None,
    x_l: Optional[np.ndarray] = None,
    x_u: Optional[np.ndarray] = None,
    u_l: Optional[np.ndarray] = None,
    u_u: Optional[np.ndarray] = None,
    verbose: bool = False,
    debug: bool = False,
    max_it: int = 100,
    time_limit: float = 1000.0,
    res_tol: float = 1e-5,
    reg_x: float = 1e0,
    reg_u: float = 1e-2,
    slew_rate: float = 0.0,
    u0_slew: Optional[np.ndarray] = None,
    lin_cost_fn: Optional[Callable] = None,
    cost_fn: Optional[Callable] = None,  # deprecated in the reference as well
    extra_cstrs_fns: Optional[Callable] = None,
    solver_settings: Optional[Dict[str, Any]] = None,
    solver_state: Optional[Dict[str, Any]] = None,
    filter_method: str = "",
    filter_window: int = 5,
    filter_it0: int = 20,
    return_min_viol: bool = False,
    min_viol_it0: int = -1,
    **extra_kw,
) -> Tuple[np.ndarray, np.ndarray, Dict[str, Any]]:
    """SCP solution of a non-linear-dynamics, quadratic-cost control problem (argument glossary:
    reference ``README.md:174-249``, ``scp_mpc.py:239-292``).  ``solver_settings["b200"]`` may carry
    the device solver's own settings (``include/pmpc_b200.h: pmpc_settings``)."""
    if cost_fn is not None:
        raise ValueError("cost_fn is deprecated, use lin_cost_fn instead.")
    t0 = time.time()
    reg_x, reg_u = float(reg_x), float(reg_u)
    single, dims, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u = _normalise(
        Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u)
    M, N, xdim, udim = dims
    slew_rate = slew_rate if slew_rate is None else float(slew_rate)
    u0_slew = np.array(u0_slew, dtype=np.float64) if u0_slew is not None else None
    solver_settings = copy(solver_settings) if solver_settings is not None else dict()

    def shape_out(X, U, data):
        if single:
            return X.reshape(N + 1, xdim), U.reshape(N, udim), data
        return X.reshape(M, N + 1, xdim), U.reshape(M, N, udim), data

    # general stage-local linear rows: static ones through solver_settings["lin_cstrs"] = (Gu, Gx, h), per-iteration ones
    # through the reference's `extra_cstrs_fns` hook (linear rows only; lin_cstrs.py)
    static_lin = lin_cstrs.normalise(solver_settings.pop("lin_cstrs", None), M, N, xdim, udim)
    has_lin = static_lin is not None or extra_cstrs_fns is not None or "extra_cstrs" in solver_settings
    device_ok = (isinstance(f_fx_fu_fn, BuiltinDynamics) and filter_method == "" and lin_cost_fn is None and not has_lin
                 and not debug and not return_min_viol and time_limit >= 1000.0
                 and not solver_settings.get("host_loop", False))
    if device_ok:
        X, U, data = _device_scp(f_fx_fu_fn, dims, Q, R, x0, X_ref, U_ref, X_prev, U_prev, x_l, x_u, u_l, u_u, max_it,
                                 res_tol, reg_x, reg_u, slew_rate, u0_slew, solver_settings, verbose, t0)
        if X is None:
            return None, None, None
        return shape_out(X, U, data)

    # ---- host loop: follows scp_mpc.py:323-442 ---------------------------------------------------
    data = dict(solver_data=[], hist=[], sol_hist=[])
    Fs = []
    field_names = ["it", "elaps", "obj", "resid", "reg_x", "reg_u"]
    tp = TablePrinter(field_names, fmts=["%04d", "%8.3e", "%8.3e", "%8.3e", "%8.3e", "%8.3e"])
    min_viol = math.inf
    if verbose:
        print_fn(tp.make_header())
    it = 0
    X = U = None
    max_res = math.inf
    while it < max_it:
        X_ = np.concatenate([x0[..., None, :], X_prev[..., :-1, :]], -2)
        f, fx, fu = f_fx_fu_fn(X_, U_prev)
        f = np.asarray(f, dtype=np.float64).reshape((M, N, xdim))
        fx = np.asarray(fx, dtype=np.float64).reshape((M, N, xdim, xdim))
        fu = np.asarray(fu, dtype=np.float64).reshape((M, N, xdim, udim))

        problems = dict(extra_kw, f_fx_fu_fn=f_fx_fu_fn, f=f, fx=fx, fu=fu, x0=x0, X_prev=X_prev, U_prev=U_prev,
                        slew_rate=slew_rate, u0_slew=u0_slew, x_l=x_l, x_u=x_u, u_l=u_l, u_u=u_u, Q=Q, R=R, X_ref=X_ref,
                        U_ref=U_ref)
        X_ref_, U_ref_ = _augment_cost(lin_cost_fn, X_prev, U_prev, Q, R, X_ref, U_ref, problems)
        solver_settings["solver_state"] = solver_state
        lin = static_lin
        if extra_cstrs_fns is not None:  # scp_mpc.py:353-354 calls it; README.md:217 says "a list of functions": both work
            fns = extra_cstrs_fns if isinstance(extra_cstrs_fns, (list, tuple)) else [extra_cstrs_fns]
            solver_settings["extra_cstrs"] = tuple(c for fn in fns for c in fn(X_prev, U_prev, problems))
        if "extra_cstrs" in solver_settings:
            dyn = lin_cstrs.from_extra_cstrs(solver_settings["extra_cstrs"], M, N, xdim, udim, solver_settings.get("Nc", -1))
            lin = dyn if lin is None else tuple(np.concatenate([a, b], 2) for a, b in zip(lin, dyn))

        t_aff = time.time()
        if lin is None:
            X, U, solver_data = aff_solve(f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref_, U_ref_, reg_x, reg_u, slew_rate,
                                          u0_slew, x_l, x_u, u_l, u_u, solver_settings=solver_settings)
        else:  # the row values ride along as extra states whose box is the constraint (lin_cstrs.py)
            fa, fxa, fua, x0a, Xpa, Qa, Xra, xla, xua = lin_cstrs.augment(lin, f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref_,
                                                                          reg_x, x_l, x_u)
            ss = {k: v for k, v in solver_settings.items() if k not in ("extra_cstrs", "lin_cstrs")}
            Xa, U, solver_data = aff_solve(fa, fxa, fua, x0a, Xpa, U_prev, Qa, R, Xra, U_ref_, reg_x, reg_u, slew_rate,
                                           u0_slew, xla, xua, u_l, u_u, solver_settings=ss)
            X = Xa[..., :xdim]
        t_aff = time.time() - t_aff
        solver_state = solver_data.get("solver_state", None)
        X, U = X.reshape((M, N + 1, xdim)), U.reshape((M, N, udim))

        if debug or filter_method != "":
            data["sol_hist"].append((X, U))
        if filter_method != "":  # mix the last iterates (scp_mpc.py:376-387)
            Xfull_prev = np.concatenate([x0[..., None, :], X_prev], -2)
            Fs.append(np.concatenate([(X - Xfull_prev).reshape(-1), (U - U_prev).reshape(-1)]))
            if it >= filter_it0:
                k = min(filter_window, len(Fs))
                alfs = FILTER_MAP[filter_method](Fs[-k:])
                XUs = data["sol_hist"][-k:]
                X = sum(a * Xi for (a, (Xi, _)) in zip(alfs, XUs))
                U = sum(a * Ui for (a, (_, Ui)) in zip(alfs, XUs))

        if np.any(np.isnan(X)) or np.any(np.isnan(U)):  # solver failure (scp_mpc.py:391-394)
            if verbose:
                print_fn("Solver failed...")
            return None, None, None

        Xn = X[..., 1:, :]
        if filter_method != "":
            dX, dU = data["sol_hist"][-1][0][..., 1:, :] - X_prev, data["sol_hist"][-1][1] - U_prev
        else:
            dX, dU = Xn - X_prev, U - U_prev
        max_res = max(np.max(np.linalg.norm(dX, 2, -1)), np.max(np.linalg.norm(dU, 2, -1)))
        dX, dU = Xn - X_ref, U - U_ref
        obj = (np.sum(dX * _bmv(Q, dX)) + np.sum(dU * _bmv(R, dU))) / N / M
        X_prev, U_prev = Xn, U

        vals = (it + 1, time.time() - t0, obj, max_res, reg_x, reg_u)
        if verbose:
            print_fn(tp.make_values(vals))
        data["solver_data"].append(solver_data)
        data["hist"].append(dict(zip(field_names, vals)))
        data.setdefault("t_aff_solve", []).append(t_aff)
        if return_min_viol and (it >= min_viol_it0 or min_viol_it0 < 0) and min_viol > max_res:
            data["min_viol_sol"], min_viol = (X, U), max_res

        if max_res < res_tol:
            break
        it += 1
        if (time.time() - t0) * (it + 1) / it > time_limit:
            break

    if verbose:
        print_fn(tp.make_footer())
        if max_res > 1e-2:
            print_fn("#" * 73)
            print_fn("Bad solution found, the solution is approximate to a residual:", "%9.4e" % max_res)
            print_fn("#" * 73)
    if not debug:
        del data["sol_hist"]
    return shape_out(X, U, data)


def solve(*args, **kwargs):
    """Alias of :func:`scp_solve` (reference ``scp_mpc.py:446-456``; ``profile=True`` is accepted and
    reports wall time per phase instead of requiring ``line_profiler``)."""
    if kwargs.pop("profile", False):
        t = time.time()
        ret = scp_solve(*args, **kwargs)
        print_fn("pmpc_b200.solve: %.3f ms total" % (1e3 * (time.time() - t)))
        if ret[2] is not None and "device" in ret[2]:
            print_fn("  device kernel: %.3f ms" % ret[2]["device"]["kernel_ms"])
        return ret
    return scp_solve(*args, **kwargs)


def tune_scp(*args, sample_nb: int = 14, reg_rng: Tuple[int, int] = (-3, 3), solve_fn: Optional[Callable] = None,
             savefig: Optional[str] = None, reg_list: Optional[Sequence[float]] = None, **kwargs):
    """Hyper-parameter sweep of the proximal weights (reference ``scp_mpc.py:460-497``): solves the problem for
    ``sample_nb`` values of ``reg_x`` on a log grid (``reg_u = reg_ratio * reg_x``) and returns the pair with the
    smallest final residual (``reg_list`` gives the grid explicitly).  A failed solve counts as residual 1e2, as in the
    reference.  The residual curve is
    plotted only when matplotlib is importable and ``savefig`` is given (the reference always plots)."""
    solve_fn = scp_solve if solve_fn is None else solve_fn
    reg_ratio = kwargs.pop("reg_ratio", 1e-1)
    reg_list = np.asarray(np.logspace(*reg_rng, sample_nb) if reg_list is None else reg_list, dtype=np.float64)
    res_list = []
    for reg in reg_list:
        kwargs.update(reg_x=float(reg), reg_u=float(reg * reg_ratio), verbose=False)
        X, U, data = solve_fn(*args, **kwargs)
        res_list.append(1e2 if data is None else data["hist"][-1]["resid"])  # (the handle is re-used through the cache)
    if savefig is not None:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt

            plt.figure()
            plt.loglog(reg_list, res_list)
            plt.ylabel("final residual")
            plt.xlabel("reg_x")
            plt.title("reg_u = reg_x * %6.1e" % reg_ratio)
            plt.grid(True, which="both")
            plt.tight_layout()
            plt.savefig(savefig, dpi=200)
        except ImportError:
            pass
    reg_x = float(reg_list[int(np.argmin(res_list))])
    return reg_x, reg_ratio * reg_x


def solve_problems(problems: List[Dict[str, Any]], verbose: bool = False, **kw):
    """Solve a list of problem dicts (reference ``scp_mpc.py:504-511`` does it serially).

    Problems that share shapes, hyper-parameters and a built-in dynamics model are batched into
    ONE device launch (one CTA per problem); everything else falls back to a loop over ``solve``.
    """
    from .batch import solve_batched

    return solve_batched(problems, verbose=verbose, **kw)

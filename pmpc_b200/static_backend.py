"""Marshalling layer between the SCP loop and the native boundary.

Mirrors reference ``pmpc/static_backend.py``: shape asserts and the Fortran re-layout of
``lqp_solve`` (:24-104), and the routing / NaN-sentinel / default logic of ``aff_solve``
(:198-312) — including its quirks: ``Nc`` defaults to -1 = full consensus
(:257), ``slew_reg`` comes from the ``slew_rate`` kwarg (:262) while ``slew_reg0`` comes from
``solver_settings["slew_reg"]`` (:263-267).

Difference, stated once: this back end implements the **QP (sum-of-costs) formulation**
(``main.jl:115-171``).  The reference's cone route (default ``solver="ecos"``) has the same
minimiser for a single particle with hard constraints; for M > 1 its epigraph objective differs
 and smoothing (``smooth_alpha``) is a different problem.  Such requests are served
by the QP formulation with a one-time warning instead of silently returning something else.
"""
from __future__ import annotations

import math
import warnings
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _lib, pmpcjl
from .utils import atleast_nd, to_numpy_f64

_CONE_SOLVERS = ("ecos", "gurobi", "mosek", "cosmo")
_warned = set()


def is_precompiled_backend_available() -> bool:
    """True when ``libpmpc_b200.so`` is built (reference: pmpcjl importable, :14-21)."""
    try:
        _lib.load()
        return True
    except Exception:
        return False


def py2jl(x: np.ndarray, keep: int = 1) -> np.ndarray:
    """``(batch..., a[, b])`` -> ``(a[, b], reversed batch...)`` view (``julia_utils.py:69-77``)."""
    n = x.ndim
    return np.transpose(x, tuple(range(n - keep, n)) + tuple(range(n - keep - 1, -1, -1)))


def _check_shapes(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, slew_reg, slew_reg0, slew_um1):
    assert x0.ndim == 2 and f.ndim == 3 and fu.ndim == 4
    xdim, M = x0.shape
    N, udim = f.shape[1], fu.shape[1]
    expect = dict(f=(xdim, N, M), fx=(xdim, xdim, N, M), fu=(xdim, udim, N, M), X_prev=(xdim, N, M), U_prev=(udim, N, M),
                  Q=(xdim, xdim, N, M), R=(udim, udim, N, M), X_ref=(xdim, N, M), U_ref=(udim, N, M), lx=(xdim, N, M),
                  ux=(xdim, N, M), lu=(udim, N, M), uu=(udim, N, M), slew_um1=(udim, M), slew_reg=(M,), slew_reg0=(M,))
    have = dict(f=f, fx=fx, fu=fu, X_prev=X_prev, U_prev=U_prev, Q=Q, R=R, X_ref=X_ref, U_ref=U_ref, lx=lx, ux=ux, lu=lu,
                uu=uu, slew_um1=slew_um1, slew_reg=slew_reg, slew_reg0=slew_reg0)
    for k, shp in expect.items():
        assert have[k].shape == shp, (k, have[k].shape, shp)
    return xdim, udim, N, M


def lqp_solve(Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u, slew_reg, slew_reg0,
              slew_um1, verbose=False):
    """Reference ``static_backend.lqp_solve`` (:24-104): Julia-layout arrays in, ``(M,N,n)`` out."""
    xdim, udim, N, M = _check_shapes(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, slew_reg,
                                     slew_reg0, slew_um1)
    X, U = pmpcjl.lqp_solve(int(xdim), int(udim), int(N), int(M), int(Nc), x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref,
                            U_ref, lx, ux, lu, uu, float(reg_x), float(reg_u), slew_reg, slew_reg0, slew_um1, int(verbose))
    return np.reshape(X, (M, N, xdim)), np.reshape(U, (M, N, udim))


def lcone_solve(Nc, x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, reg_x, reg_u, slew_reg, slew_reg0,
                slew_um1, smooth_alpha=1e1, verbose=False, solver="ecos"):
    """Reference ``static_backend.lcone_solve`` (:107-191)."""
    xdim, udim, N, M = _check_shapes(x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref, U_ref, lx, ux, lu, uu, slew_reg,
                                     slew_reg0, slew_um1)
    X, U = pmpcjl.lcone_solve(int(xdim), int(udim), int(N), int(M), int(Nc), x0, f, fx, fu, X_prev, U_prev, Q, R, X_ref,
                              U_ref, lx, ux, lu, uu, float(reg_x), float(reg_u), slew_reg, slew_reg0, slew_um1,
                              int(verbose), float(smooth_alpha), solver)
    return np.reshape(X, (M, N, xdim)), np.reshape(U, (M, N, udim))


def aff_solve(f, fx, fu, x0, X_prev, U_prev, Q, R, X_ref, U_ref, reg_x, reg_u, slew_rate, u_slew, x_l, x_u, u_l, u_u,
              solver_settings: Optional[Dict[str, Any]] = None) -> Tuple[np.ndarray, np.ndarray, Any]:
    """Solve one linearised MPC sub-problem (reference ``static_backend.aff_solve`` :198-312)."""
    solver_settings = dict(solver_settings) if solver_settings is not None else dict()
    f = atleast_nd(to_numpy_f64(f), 3)
    fx, fu = atleast_nd(to_numpy_f64(fx), 4), atleast_nd(to_numpy_f64(fu), 4)
    x0 = atleast_nd(to_numpy_f64(x0), 2)
    X_prev, U_prev = atleast_nd(to_numpy_f64(X_prev), 3), atleast_nd(to_numpy_f64(U_prev), 3)
    Q, R = atleast_nd(to_numpy_f64(Q), 4), atleast_nd(to_numpy_f64(R), 4)
    X_ref, U_ref = atleast_nd(to_numpy_f64(X_ref), 3), atleast_nd(to_numpy_f64(U_ref), 3)
    M = x0.shape[0]
    bounds = [None if (z is None or np.size(z) == 0) else atleast_nd(to_numpy_f64(z), 3) for z in (x_l, x_u, u_l, u_u)]

    solver = str(solver_settings.get("solver", "ecos")).lower()
    assert solver in _CONE_SOLVERS + ("osqp",), solver
    cone_requested = solver in _CONE_SOLVERS or "smooth_cstr" in solver_settings or "smooth_alpha" in solver_settings
    if cone_requested:
        smoothing = (not math.isnan(float(solver_settings.get("smooth_alpha", math.nan)))) or "smooth_cstr" in solver_settings
        if (M > 1 or smoothing) and "cone" not in _warned:
            _warned.add("cone")
            warnings.warn("pmpc_b200 solves the QP (sum-of-costs, hard box) formulation of lqp_solve; the requested "
                          "cone/smoothed route is served by it (see DESIGN.md, 'out of scope').")

    Nc = solver_settings.get("Nc", -1)
    nan_like = lambda z: np.full_like(z, np.nan)
    Xj, Uj = py2jl(X_prev, 1), py2jl(U_prev, 1)
    x_l, x_u = [nan_like(Xj) if b is None else py2jl(b, 1) for b in bounds[:2]]
    u_l, u_u = [nan_like(Uj) if b is None else py2jl(b, 1) for b in bounds[2:]]
    slew_reg = np.full((M,), math.nan if slew_rate is None else float(slew_rate))
    slew_reg0 = np.full((M,), float(solver_settings["slew_reg"]) if "slew_reg" in solver_settings else math.nan)
    slew_um1 = nan_like(Uj[:, 0, :]) if u_slew is None else py2jl(atleast_nd(to_numpy_f64(u_slew), 2), 1)
    X, U = lqp_solve(Nc, py2jl(x0, 1), py2jl(f, 1), py2jl(fx, 2), py2jl(fu, 2), Xj, Uj, py2jl(Q, 2), py2jl(R, 2),
                     py2jl(X_ref, 1), py2jl(U_ref, 1), x_l, x_u, u_l, u_u, reg_x, reg_u, slew_reg, slew_reg0, slew_um1,
                     verbose=bool(solver_settings.get("verbose", False)))
    X_traj = np.concatenate([x0[:, None, :], X], -2)
    return X_traj, U, dict()

"""pmpc_b200 — B200-native drop-in for the ``solve`` path of StanfordASL/pmpc.

``from pmpc_b200 import solve, Problem`` mirrors ``from pmpc import solve, Problem``
(reference ``pmpc/__init__.py:1-35``).  All numerical work runs in ``libpmpc_b200.so``
(hand-written sm_100a CUDA, C ABI in ``include/pmpc_b200.h``); there is no CPU fallback.
"""
from .scp_mpc import (  # noqa: F401
    AA_method, FILTER_MAP, scp_solve, select_method, smooth_method, solve, solve_problems, tune_scp,
)
from .accelerated import accelerated_scp_solve  # noqa: F401
from .static_backend import aff_solve, is_precompiled_backend_available, lqp_solve  # noqa: F401
from .dynamics import BuiltinDynamics, DubinsCar, LTI, LTV, UserDynamics  # noqa: F401
from .handle import DeviceSolver  # noqa: F401
from .batch import solve_batch  # noqa: F401
from ._lib import PmpcError, pinned_copy, pinned_empty  # noqa: F401
from .problem_struct import Problem  # noqa: F401
from .lqp_matrices import lqp_generate_problem_matrices  # noqa: F401

"""Keyword-compatible arguments in the solver scp solve call (reference pmpc/__init__.py:4-33)."""
SOLVE_KWS = {
    "X_ref", "U_ref", "X_prev", "U_prev", "x_l", "x_u", "u_l", "u_u", "verbose", "debug", "max_it", "time_limit",
    "res_tol", "reg_x", "reg_u", "slew_rate", "u_slew", "cost_fn", "extra_cstrs_fns", "method", "solver_settings",
    "solver_state", "filter_method", "filter_window", "filter_it0",
}

__version__ = "0.1.0"

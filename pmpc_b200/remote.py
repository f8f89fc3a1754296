"""Client-side surface of the reference's task farm (``pmpc/remote.py``), served in-process by the GPU.

The reference fans independent problems out to worker processes over ZeroMQ (worker discovery through redis) because
one Julia/OSQP solve occupies one CPU core (``remote.py:391-452``).  Here a list of problems is ONE batched launch
(one CTA per problem, ``pmpc_b200.solve_problems``), so the farm is unnecessary: this module keeps the functions a
client of ``pmpc.remote`` calls — same names, arguments and return values — and runs them locally.  The wire protocol
(ZeroMQ + cloudpickle + zstd) and the server / redis registry are NOT provided (out of scope); the
server-management functions are accepted and do nothing, so that scripts written against the reference run unchanged.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from . import scp_mpc

DEFAULT_PORT = 65535 - 9117  # reference remote.py default; kept for scripts that pass it around
_POSITIONAL = ("f_fx_fu_fn", "Q", "R", "x0")


def solve(*args, **kw):
    """``pmpc.remote.solve`` (``remote.py:100-106``): the reference forwards to a worker; here it is the local call."""
    return scp_mpc.solve(*args, **kw)


def tune_scp(*args, **kw):
    return scp_mpc.tune_scp(*args, **kw)


for _f in (solve, tune_scp):  # attributes clients set on the reference's functions to pick a worker
    _f.hostname, _f.port, _f.blocking = "localhost", DEFAULT_PORT, True


def solve_problem(problem: Dict[str, Any], solve_fn: Callable = solve) -> Tuple[np.ndarray, np.ndarray, Dict[str, Any]]:
    """One problem given as a keyword dictionary (``remote.py:324-343``): ``f_fx_fu_fn, Q, R, x0`` are passed
    positionally, everything else as keywords; ``verbose`` defaults to True as in the reference."""
    rest = {k: v for k, v in problem.items() if k not in _POSITIONAL}
    rest.setdefault("verbose", True)
    if "extra_cstrs_fn_np" in rest:
        rest["extra_cstrs_fn"] = rest["extra_cstrs_fn_np"]
    return solve_fn(*(problem[k] for k in _POSITIONAL), **rest)


def solve_problems(problems: List[Optional[Dict[str, Any]]], verbose: bool = False, randomize_assignment: bool = True,
                   rescan: bool = True, max_solve_time: float = 20.0, redis_config: Optional[Dict[str, Any]] = None):
    """``pmpc.remote.solve_problems`` (``remote.py:391-452``): solutions in the order of ``problems``; a ``None`` entry
    yields ``None``.  Worker-pool arguments are accepted and ignored: same-shaped problems with device dynamics go to the
    GPU as one batch, the rest are solved one after the other."""
    todo = [(i, dict(p, verbose=False)) for i, p in enumerate(problems) if p is not None]
    out: List[Any] = [None] * len(problems)
    for (i, _), res in zip(todo, scp_mpc.solve_problems([p for _, p in todo], verbose=verbose)):
        out[i] = res
    return out


def start_server(port: int = DEFAULT_PORT, verbose: bool = False, redis_config: Optional[Dict[str, Any]] = None):
    """No worker processes are needed; returns immediately (``remote.py:123-130`` starts a ZeroMQ server there)."""
    if verbose:
        print("pmpc_b200.remote: problems are solved in-process on the GPU; no server started (port %d ignored)" % port)


def precompilation_call():
    """The reference warms the Julia back end up with a tiny solve (``remote.py:133-169``); the CUDA library needs no
    warm-up beyond loading, which this does (and fails loudly without a device or the library)."""
    from . import _lib

    lib = _lib.load()
    if lib.pmpc_device_count() < 1:
        raise _lib.PmpcError("pmpc_b200.remote.precompilation_call: no CUDA device (there is no CPU fallback)")

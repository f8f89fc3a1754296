"""Multi-GPU plumbing: one process per GPU (``torch.distributed``), problems sharded, no data-path
collective (SURVEY.md 8e case 1: independent problems shard naturally).

The reference's counterpart is the ZeroMQ task farm ``pmpc/remote.py:391-452`` (one problem per CPU
worker process over TCP).  Here rank r owns the contiguous slice ``shard_bounds(B, r, world)`` of the
problem batch and runs it on its own B200; only the per-rank wall/device time is reduced (max) for
reporting.  
Particle sharding of ONE consensus problem (8e case 2, BASELINE config 5): rank r holds the particles
``shard_bounds(M, r, world)``; the Nc*m consensus terms and the pass statistics are all-reduced INSIDE
the solve kernel through peer-mapped exchange buffers (``Solver::xreduce`` in csrc/solver.cuh), so
the host only has to all-gather the 64-byte CUDA IPC handles once (``solve_consensus_sharded``).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Balanced contiguous partition: the first ``total % world`` ranks get one extra item."""
    assert 0 <= rank < world and total >= 0
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: str = "nccl"):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world == 1:
        return None
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def reduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks of a scalar (timing): the job finishes when the slowest rank does."""
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, dist=None, device=None) -> float:
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj, dist=None) -> List:
    """All ranks' python objects on every rank, in rank order (results of sharded batches)."""
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def solve_problems_sharded(problems: Sequence[dict], dist=None, device: int = -1):
    """Rank-local slice of ``problems`` through the batched device path; every rank returns the full,
    ordered result list (gathered)."""
    from .batch import solve_batched

    rank, world = (0, 1) if dist is None else (dist.get_rank(), dist.get_world_size())
    lo, hi = shard_bounds(len(problems), rank, world)
    local = solve_batched(list(problems[lo:hi]), device=device)
    parts = gather_objects(local, dist)
    return [r for part in parts for r in part]


def shard_particles(arrs: dict, lo: int, hi: int, Nc: int) -> dict:
    """Rank-local slice ``[lo, hi)`` of the particle axis (axis 0) of every array of a consensus problem.
    The shared block's bounds belong to global particle 0 (``lqp_utils.jl:329-330``): they are copied
    into the local particle 0 so that every rank projects the shared controls identically."""
    import numpy as np

    out = {}
    for k, v in arrs.items():
        if v is None:
            out[k] = None
            continue
        a = np.array(np.asarray(v, dtype=np.float64)[lo:hi])
        if k in ("u_l", "u_u") and Nc > 0:
            a[0, :Nc] = np.asarray(v, dtype=np.float64)[0, :Nc]
        out[k] = a
    return out


def _shard_solver(rank, world, dyn_params, full, M, N, n, m, Nc, reg_x, reg_u, slew_rate, dyn_eps, device, ltv, settings):
    """Handle of rank `rank` of `world`: its slice of the particles uploaded, dynamics set, not yet connected."""
    import numpy as np

    from .handle import DeviceSolver

    Nc_eff = N if Nc < 0 else int(Nc)
    lo, hi = shard_bounds(M, rank, world)
    assert hi - lo >= 2, "particle sharding needs at least 2 particles per rank"
    loc = shard_particles(full, lo, hi, Nc_eff)
    ds = DeviceSolver(hi - lo, N, n, m, Nc, B=1, device=device, settings=settings)
    ds.set_problem(loc["x0"], loc["Q"], loc["R"], loc["X_ref"], loc["U_ref"], None, None, loc["u_l"], loc["u_u"], reg_x,
                   reg_u, None if slew_rate is None else np.full(hi - lo, float(slew_rate)))
    ds.set_warm_start(loc["X_ref"] if loc["X_prev"] is None else loc["X_prev"],
                      loc["U_ref"] if loc["U_prev"] is None else loc["U_prev"])
    if ltv is not None:
        A, Bm, c = ltv
        ds.set_ltv(np.asarray(A)[lo:hi], np.asarray(Bm)[lo:hi], None if c is None else np.asarray(c)[lo:hi])
    else:
        ds.set_dubins(np.asarray(dyn_params, dtype=np.float64)[lo:hi], dyn_eps)
    return ds


def solve_consensus_sharded(dyn_params, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None, u_l=None, u_u=None,
                            max_it: int = 100, res_tol: float = 1e-5, reg_x: float = 1.0, reg_u: float = 1e-2,
                            slew_rate=0.0, Nc: int = 1, dyn_eps: float = 1e-6, dist=None, device: int = -1,
                            ltv=None, settings=None):
    """ONE consensus problem (arrays with leading particle axis M, as ``pmpc.solve`` takes them) with its
    particles sharded over the ranks of ``dist`` (one process per GPU of one node).  Dynamics: Dubins car
    with per-particle parameters ``dyn_params (M, 3)``, or ``ltv=(A, B, c)`` with ``A (M,N,n,n)``.
    Every rank passes the same full arrays and gets the same full result ``X (M,N+1,n), U (M,N,m), info``.
    Control box only (a state box needs the unsharded solver)."""
    import numpy as np

    rank, world = (0, 1) if dist is None else (dist.get_rank(), dist.get_world_size())
    Q = np.asarray(Q, dtype=np.float64)
    M, N, n = Q.shape[:3]
    m = np.asarray(R).shape[-1]
    full = dict(Q=Q, R=R, x0=x0, X_ref=X_ref, U_ref=U_ref, X_prev=X_prev, U_prev=U_prev, u_l=u_l, u_u=u_u)
    ds = _shard_solver(rank, world, dyn_params, full, M, N, n, m, Nc, reg_x, reg_u, slew_rate, dyn_eps, device, ltv, settings)
    if world > 1:
        handles = gather_objects(ds.comm_export(), dist)
        ds.comm_connect(rank, world, handles, M)
        dist.barrier()  # every peer buffer is mapped before any kernel spins on it
    ds.scp_solve(max_it, res_tol)
    X, U = ds.trajectory()
    info = {k: int(v[0]) for k, v in ds.info().items()}
    info["kernel_ms"] = ds.last_kernel_ms()
    parts = gather_objects((X[0], U[0]), dist)
    if world > 1:
        dist.barrier()  # nobody unmaps while a peer may still read
    return np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts], 0), info


def solve_consensus_sharded_local(world: int, dyn_params, Q, R, x0, X_ref=None, U_ref=None, X_prev=None, U_prev=None,
                                  u_l=None, u_u=None, max_it: int = 100, res_tol: float = 1e-5, reg_x: float = 1.0,
                                  reg_u: float = 1e-2, slew_rate=0.0, Nc: int = 1, dyn_eps: float = 1e-6, devices=None,
                                  ltv=None, settings=None, timeout_s=None):
    """The same particle sharding driven from ONE process: ``world`` handles, handle r on ``devices[r]`` (default: all on
    the current device, where their kernels run concurrently on the handles' own streams and exchange the consensus
    terms through the same in-kernel protocol, plain device pointers instead of IPC mappings).  All solves are launched
    asynchronously before any is waited for.  Returns ``X (M,N+1,n), U (M,N,m), infos`` (one info dict per rank)."""
    import numpy as np

    Q = np.asarray(Q, dtype=np.float64)
    M, N, n = Q.shape[:3]
    m = np.asarray(R).shape[-1]
    devices = [-1] * world if devices is None else list(devices)
    full = dict(Q=Q, R=R, x0=x0, X_ref=X_ref, U_ref=U_ref, X_prev=X_prev, U_prev=U_prev, u_l=u_l, u_u=u_u)
    solvers = [_shard_solver(r, world, dyn_params, full, M, N, n, m, Nc, reg_x, reg_u, slew_rate, dyn_eps, devices[r], ltv,
                             settings) for r in range(world)]
    try:
        if world > 1:
            for ds in solvers:
                ds.comm_export()
            for r, ds in enumerate(solvers):
                ds.comm_connect_local(r, solvers, M)
                if timeout_s is not None:
                    ds.comm_set_timeout(timeout_s)
        for ds in solvers:
            ds.scp_solve_async(max_it, res_tol)
        for ds in solvers:
            ds.sync()
        parts, infos = [], []
        for ds in solvers:
            X, U = ds.trajectory()
            parts.append((X[0], U[0]))
            info = {k: int(v[0]) for k, v in ds.info().items()}
            info["kernel_ms"] = ds.last_kernel_ms()
            infos.append(info)
    finally:
        for ds in solvers:
            ds.close()
    return np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts], 0), infos

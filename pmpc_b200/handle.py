"""Python face of the handle API (``pmpc_create`` ... ``pmpc_destroy``, include/pmpc_b200.h).

Arrays use the Python layout of ``pmpc.solve`` with an optional leading batch axis:
vectors ``([B,] M, N, d)``, matrices ``([B,] M, N, rows, cols)``.  They are re-laid out to the
reference C-ABI memory order (column-major blocks, ``c_interface.jl:28-46``) before the call,
exactly what ``static_backend.lqp_solve`` does with ``py2jl`` + ``asfortranarray``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check, dptr, f64c, pinned_empty

INFO_FIELDS = ("scp_iters", "admm_iters", "factorizations", "polish_passes", "status", "qp_by_polish", "qp_by_admm",
               "sweeps")


class DeviceSolver:
    def __init__(self, M: int, N: int, xdim: int, udim: int, Nc: int = -1, B: int = 1, device: int = -1,
                 settings: Optional[dict] = None):
        self.lib = _lib.load()
        self.B, self.M, self.N, self.n, self.m = int(B), int(M), int(N), int(xdim), int(udim)
        self.Nc = int(Nc)
        self._h = C.c_void_p()
        check(self.lib.pmpc_create(C.byref(self._h), device, self.B, self.M, self.N, self.n, self.m, self.Nc), "create")
        # NumPy C-order (..., rows, cols) matrix blocks are consumed as they are: the re-layout of
        # static_backend.py:71-74 (py2jl + asfortranarray) happens in the device packing kernels
        check(self.lib.pmpc_set_layout(self._h, 1), "set_layout")
        self._out = None  # page-locked result buffers, allocated on first use
        if settings:
            self.set_settings(**settings)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers -------------------------------------------------------------------------------
    def _vec(self, a, d):
        if a is None:
            return None
        a = f64c(a)
        assert a.size == self.B * self.M * self.N * d, (a.shape, (self.B, self.M, self.N, d))
        return a.reshape(self.B * self.M, self.N, d)

    def _mat(self, a, r, c):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        assert a.size == self.B * self.M * self.N * r * c, (a.shape, (self.B, self.M, self.N, r, c))
        return f64c(a).reshape(self.B * self.M, self.N, r, c)

    # inputs that are identical for every particle (of every problem) may be given once: ``(N, d)`` / ``(N, r, c)`` for the
    # whole batch, or one block per problem ``(B, N, ...)``; the device packing kernels expand them (pmpc_set_broadcast)
    _BC_BITS = dict(Q=0, R=1, X_ref=2, U_ref=3, x_l=4, x_u=5, u_l=6, u_u=7, X_prev=8, U_prev=9)

    def _bcast(self, a, tail, key, masks):
        """C-contiguous float64 view of input ``key`` and its broadcast level recorded in ``masks = [all, per_problem]``."""
        if a is None:
            return None
        a = f64c(a)
        per = self.N * int(np.prod(tail))
        if a.size == self.B * self.M * per:
            return a.reshape((self.B * self.M, self.N) + tuple(tail))
        if a.size == per:
            masks[0] |= 1 << self._BC_BITS[key]
            return a.reshape((1, self.N) + tuple(tail))
        if a.size == self.B * per:
            masks[1] |= 1 << self._BC_BITS[key]
            return a.reshape((self.B, self.N) + tuple(tail))
        raise AssertionError((key, a.shape, (self.B, self.M, self.N) + tuple(tail)))

    def _per_particle(self, a, d=None):
        if a is None:
            return None
        a = f64c(a)
        P = self.B * self.M
        if d is None:
            return np.ascontiguousarray(np.broadcast_to(a.reshape(-1) if a.size > 1 else a.reshape(1), (P,)))
        assert a.size == P * d, (a.shape, P, d)
        return a.reshape(P, d)

    # -- configuration -------------------------------------------------------------------------
    def set_settings(self, **kw):
        s = _lib.default_settings(**kw)
        check(self.lib.pmpc_set_settings(self._h, C.byref(s)), "set_settings")
        self.settings = s

    def set_problem(self, x0, Q, R, X_ref=None, U_ref=None, x_l=None, x_u=None, u_l=None, u_u=None, reg_x=1.0,
                    reg_u=1e-2, slew_reg=None, slew_reg0=None, slew_um1=None):
        n, m = self.n, self.m
        absent = lambda z: z is None or np.size(z) == 0
        masks = [0, 0]
        x0 = f64c(np.asarray(x0, float))
        if x0.size == self.B * n and self.M > 1:  # one initial state per problem: every particle starts there
            x0 = np.repeat(x0.reshape(self.B, 1, n), self.M, axis=1)
        keep = [  # keep references alive for the duration of the call
            f64c(x0.reshape(self.B * self.M, n)), self._bcast(Q, (n, n), "Q", masks), self._bcast(R, (m, m), "R", masks),
            self._bcast(X_ref, (n,), "X_ref", masks), self._bcast(U_ref, (m,), "U_ref", masks),
            None if absent(x_l) else self._bcast(x_l, (n,), "x_l", masks), None if absent(x_u) else self._bcast(x_u, (n,), "x_u", masks),
            None if absent(u_l) else self._bcast(u_l, (m,), "u_l", masks), None if absent(u_u) else self._bcast(u_u, (m,), "u_u", masks),
        ]
        sl = [self._per_particle(slew_reg), self._per_particle(slew_reg0), self._per_particle(slew_um1, m)]
        self._masks = masks
        check(self.lib.pmpc_set_broadcast(self._h, masks[0], masks[1]), "set_broadcast")
        check(self.lib.pmpc_set_problem(self._h, *[dptr(a) for a in keep], float(reg_x), float(reg_u),
                                        *[dptr(a) for a in sl]), "set_problem")

    def set_warm_start(self, X_prev=None, U_prev=None):
        masks = list(getattr(self, "_masks", [0, 0]))
        for k in ("X_prev", "U_prev"):
            masks[0] &= ~(1 << self._BC_BITS[k])
            masks[1] &= ~(1 << self._BC_BITS[k])
        a, b = self._bcast(X_prev, (self.n,), "X_prev", masks), self._bcast(U_prev, (self.m,), "U_prev", masks)
        self._masks = masks
        check(self.lib.pmpc_set_broadcast(self._h, masks[0], masks[1]), "set_broadcast")
        check(self.lib.pmpc_set_warm_start(self._h, dptr(a), dptr(b)), "set_warm_start")

    def set_dubins(self, p, eps=1e-6):
        """p: (3,) | (M,3) | (B,M,3) = (v_scale, w_scale, T) — reference tests/dubins_car.py:60."""
        p = np.asarray(p, dtype=np.float64)
        p = np.ascontiguousarray(np.broadcast_to(p, (self.B, self.M, 3)) if p.ndim < 3 else p).reshape(self.B * self.M, 3)
        check(self.lib.pmpc_set_dynamics(self._h, 1, dptr(p), 3, float(eps)), "set_dynamics")
        self._user_key = None

    def register_dynamics(self, cuda_src: str, fn_name: str, params=None):
        """User device dynamics (`pmpc_register_dynamics`): CUDA C++ source of
        ``__device__ void fn_name(const double* x, const double* u, const double* p, double* f, double* fx, double* fu)``
        compiled with NVRTC; ``params``: per-particle parameters ``(P,) | (M,P) | (B,M,P)`` or None."""
        if params is None:
            pp, npar = None, 0
        else:
            pp = np.asarray(params, dtype=np.float64)
            npar = pp.shape[-1]
            pp = np.ascontiguousarray(np.broadcast_to(pp, (self.B, self.M, npar)) if pp.ndim < 3 else pp).reshape(self.B * self.M, npar)
        key = (cuda_src, fn_name, None if pp is None else pp.tobytes())
        if getattr(self, "_user_key", None) == key:
            return  # already compiled and loaded for exactly these parameters (NVRTC costs ~0.3 s)
        check(self.lib.pmpc_register_dynamics(self._h, cuda_src.encode(), fn_name.encode(), dptr(pp), npar), "register_dynamics")
        self._user_key = key

    def set_linearization(self, f, fx, fu):
        n, m = self.n, self.m
        a, b, c = self._vec(f, n), self._mat(fx, n, n), self._mat(fu, n, m)
        check(self.lib.pmpc_set_linearization(self._h, dptr(a), dptr(b), dptr(c)), "set_linearization")
        self._user_key = None

    def set_ltv(self, A, Bm, c=None):
        n, m = self.n, self.m
        a, b, cc = self._mat(A, n, n), self._mat(Bm, n, m), self._vec(c, n)
        check(self.lib.pmpc_set_ltv(self._h, dptr(a), dptr(b), dptr(cc)), "set_ltv")
        self._user_key = None

    # -- compute -------------------------------------------------------------------------------
    def linearize(self):
        """Device linearisation (kernel K1) about the handle's (X_prev, U_prev): f, fx, fu."""
        check(self.lib.pmpc_linearize(self._h), "linearize")
        P, N, n, m = self.B * self.M, self.N, self.n, self.m
        f, fx, fu = np.empty((P, N, n)), np.empty((P, N, n, n)), np.empty((P, N, n, m))
        check(self.lib.pmpc_get_linearization(self._h, dptr(f), dptr(fx), dptr(fu)), "get_linearization")
        shp = (self.B, self.M, N) if self.B > 1 else (self.M, N)
        return f.reshape(shp + (n,)), fx.reshape(shp + (n, n)), fu.reshape(shp + (n, m))

    def qp_solve(self, keep_prev=False):
        check(self.lib.pmpc_qp_solve(self._h, int(keep_prev)), "qp_solve")

    def scp_solve(self, max_it=100, res_tol=1e-5):
        check(self.lib.pmpc_scp_solve(self._h, int(max_it), float(res_tol)), "scp_solve")

    def scp_solve_async(self, max_it=100, res_tol=1e-5):
        check(self.lib.pmpc_scp_solve_async(self._h, int(max_it), float(res_tol)), "scp_solve_async")

    def sync(self):
        check(self.lib.pmpc_sync(self._h), "sync")

    # -- results -------------------------------------------------------------------------------
    def solution(self):
        P, N = self.B * self.M, self.N
        X, U = np.empty((P, N, self.n)), np.empty((P, N, self.m))
        check(self.lib.pmpc_get_solution(self._h, dptr(X), dptr(U)), "get_solution")
        shp = (self.B, self.M, N) if self.B > 1 else (self.M, N)
        return X.reshape(shp + (self.n,)), U.reshape(shp + (self.m,))

    def trajectory(self, reuse_buffers: bool = False, out=None):
        """``X (B, M, N+1, n)`` including x0 and ``U (B, M, N, m)`` — the arrays ``solve`` returns
        (``static_backend.py:311``), assembled on the device.  With ``reuse_buffers`` the results land in
        page-locked buffers owned by this solver (overwritten by the next call) instead of fresh arrays;
        ``out=(X, U)`` writes into caller-provided C-contiguous arrays (e.g. slices of one page-locked buffer)."""
        P, N = self.B * self.M, self.N
        if out is not None:
            X, U = out
            assert X.flags.c_contiguous and U.flags.c_contiguous and X.size == P * (N + 1) * self.n and U.size == P * N * self.m
        elif reuse_buffers:
            if self._out is None:
                self._out = (pinned_empty((P, N + 1, self.n)), pinned_empty((P, N, self.m)))
            X, U = self._out
        else:
            X, U = np.empty((P, N + 1, self.n)), np.empty((P, N, self.m))
        check(self.lib.pmpc_get_trajectory(self._h, dptr(X), dptr(U)), "get_trajectory")
        return X.reshape(self.B, self.M, N + 1, self.n), U.reshape(self.B, self.M, N, self.m)

    def info(self):
        buf = np.zeros((self.B, 8), dtype=np.int64)
        check(self.lib.pmpc_get_info(self._h, buf.ctypes.data_as(_lib.c_ll_p)), "get_info")
        return {k: buf[:, i].copy() for i, k in enumerate(INFO_FIELDS)}

    def hist(self, max_it):
        buf = np.empty((4, int(max_it), self.B))
        check(self.lib.pmpc_get_hist(self._h, dptr(buf), int(max_it)), "get_hist")
        return dict(obj=buf[0], resid=buf[1], admm_iters=buf[2], kkt=buf[3])

    def occupancy(self):
        """(resident CTAs per SM, SM count) of the solve kernel this handle launches."""
        a, b = C.c_int(), C.c_int()
        check(self.lib.pmpc_get_occupancy(self._h, C.byref(a), C.byref(b)), "get_occupancy")
        return int(a.value), int(b.value)

    # -- particle sharding over the GPUs of one node ----------------------------------------------
    def comm_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange buffer (``pmpc_comm_export``)."""
        buf = C.create_string_buffer(64)
        check(self.lib.pmpc_comm_export(self._h, buf), "comm_export")
        return buf.raw

    def comm_connect(self, rank: int, world: int, handles, M_total: int):
        """Map the peers' exchange buffers; ``handles`` = the ``world`` IPC handles in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 64 * world
        check(self.lib.pmpc_comm_connect(self._h, int(rank), int(world), C.c_char_p(blob), int(M_total)), "comm_connect")

    def comm_connect_local(self, rank: int, solvers, M_total: int):
        """Particle sharding between handles of THIS process (``pmpc_comm_connect_local``): ``solvers`` are the
        ``world`` :class:`DeviceSolver` objects in rank order, ``solvers[rank] is self``; every one must have called
        :meth:`comm_export` (which allocates its exchange buffer)."""
        arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
        check(self.lib.pmpc_comm_connect_local(self._h, int(rank), len(solvers), arr, int(M_total)), "comm_connect_local")

    def comm_set_timeout(self, seconds: float):
        check(self.lib.pmpc_comm_set_timeout(self._h, float(seconds)), "comm_set_timeout")

    def timer_start(self):
        """CUDA event on the handle's stream (the stream every kernel of this handle runs on)."""
        check(self.lib.pmpc_timer_start(self._h), "timer_start")

    def timer_stop(self) -> float:
        """Milliseconds of device time since :meth:`timer_start` (records, synchronises)."""
        ms = C.c_float()
        check(self.lib.pmpc_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return float(ms.value)

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.pmpc_last_kernel_ms(self._h, C.byref(ms)), "last_kernel_ms")
        return float(ms.value)

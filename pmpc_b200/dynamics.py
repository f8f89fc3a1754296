"""Built-in dynamics whose linearisation runs on the device (kernel K1).

An instance is a drop-in ``f_fx_fu_fn``: calling it with ``(X, U)`` returns ``(f, fx, fu)``
computed by the CUDA linearisation kernel (so it also works inside the reference-style host
loop), and ``solve`` recognises it and runs the whole SCP loop on the device instead.
"""
from __future__ import annotations

import numpy as np


class BuiltinDynamics:
    kind = "affine"

    def __getstate__(self):
        """A model is its description: the device handle a call may have cached stays behind when it is pickled
        (problems travel to ``pmpc_b200.remote`` workers as pickles)."""
        state = dict(self.__dict__)
        if "_solver" in state:
            state["_solver"] = None
        return state


class DubinsCar(BuiltinDynamics):
    """Unicycle car of reference ``tests/dubins_car.py:48-90``; ``p = (v_scale, w_scale, T)`` of
    shape ``(3,)`` or ``(M, 3)`` (perturbed-dynamics particles, ``tests/root_testing.ipynb`` cell 10),
    ``eps`` the +-eps control offset (1e-6 in ``tests/``, 1e-3 in the reference's GPU notebooks)."""

    kind = "dubins"
    xdim, udim = 4, 2

    def __init__(self, p=(1.0, 1.0, 0.3), eps: float = 1e-6):
        self.p = np.asarray(p, dtype=np.float64)
        self.eps = float(eps)
        self._solver = None

    def params_for(self, M: int) -> np.ndarray:
        return np.ascontiguousarray(np.broadcast_to(self.p, (M, 3)))

    def __call__(self, X, U):
        """(M,N,4),(M,N,2) -> f (M,N,4), fx (M,N,4,4), fu (M,N,4,2); X[:, j] is the state the
        control U[:, j] is applied to (scp_mpc.py:338-342)."""
        from .handle import DeviceSolver

        X, U = np.asarray(X, np.float64), np.asarray(U, np.float64)
        single = X.ndim == 2
        if single:
            X, U = X[None], U[None]
        M, N = X.shape[:2]
        s = self._solver
        if s is None or (s.M, s.N) != (M, N):
            s = self._solver = DeviceSolver(M, N, 4, 2, 0)
            s.set_dubins(self.params_for(M), self.eps)
        # the kernel linearises at X_ = [x0, X_prev[:-1]]: feed x0 = X[:, 0], X_prev[:-1] = X[:, 1:]
        eye4, eye2 = np.broadcast_to(np.eye(4), (M, N, 4, 4)), np.broadcast_to(np.eye(2), (M, N, 2, 2))
        s.set_problem(X[:, 0], eye4, eye2)
        s.set_warm_start(np.concatenate([X[:, 1:], X[:, -1:]], 1), U)
        f, fx, fu = s.linearize()
        return (f[0], fx[0], fu[0]) if single else (f, fx, fu)


class UserDynamics(BuiltinDynamics):
    """User device dynamics: CUDA C++ source of one ``__device__`` function

        __device__ void NAME(const double* x, const double* u, const double* p, double* f, double* fx, double* fu)

    (``x`` the state ``u`` is applied to, ``p`` this particle's parameters; writes the next state ``f``, row-major
    ``fx[r*xdim+t] = df_r/dx_t`` and ``fu[r*udim+t] = df_r/du_t``).  It replaces the Python ``f_fx_fu_fn`` callable of
    the reference (``pmpc/scp_mpc.py:338-342``): the library compiles it with NVRTC for sm_100a
    (``pmpc_register_dynamics``) and the whole SCP loop stays on the device.  An instance is also callable like an
    ordinary ``f_fx_fu_fn`` (it runs the compiled linearisation kernel), so it works in the host loop too.
    ``params``: ``(P,)`` or ``(M, P)`` per-particle parameters, or None."""

    kind = "user"

    def __init__(self, cuda_src: str, fn_name: str, xdim: int, udim: int, params=None):
        self.src, self.fn_name, self.xdim, self.udim = cuda_src, fn_name, int(xdim), int(udim)
        self.params = None if params is None else np.asarray(params, dtype=np.float64)
        self._solver = None

    def params_for(self, M: int):
        if self.params is None:
            return None
        return np.ascontiguousarray(np.broadcast_to(self.params, (M, self.params.shape[-1])))

    def __call__(self, X, U):
        from .handle import DeviceSolver

        X, U = np.asarray(X, np.float64), np.asarray(U, np.float64)
        single = X.ndim == 2
        if single:
            X, U = X[None], U[None]
        M, N = X.shape[:2]
        n, m = self.xdim, self.udim
        s = self._solver
        if s is None or (s.M, s.N) != (M, N):
            s = self._solver = DeviceSolver(M, N, n, m, 0)
            s.register_dynamics(self.src, self.fn_name, self.params_for(M))
        eye_n, eye_m = np.broadcast_to(np.eye(n), (M, N, n, n)), np.broadcast_to(np.eye(m), (M, N, m, m))
        s.set_problem(X[:, 0], eye_n, eye_m)
        s.set_warm_start(np.concatenate([X[:, 1:], X[:, -1:]], 1), U)
        f, fx, fu = s.linearize()
        return (f[0], fx[0], fu[0]) if single else (f, fx, fu)


class LTV(BuiltinDynamics):
    """x+ = A_j x + B_j u + c_j with per-(particle, stage) matrices ``A (M,N,n,n)``, ``B (M,N,n,m)``,
    ``c (M,N,n)`` or None.  Linear, so its linearisation is itself."""

    kind = "ltv"

    def __init__(self, A, B, c=None):
        self.A, self.B = np.asarray(A, np.float64), np.asarray(B, np.float64)
        self.c = None if c is None else np.asarray(c, np.float64)
        self.xdim, self.udim = self.B.shape[-2], self.B.shape[-1]

    def __call__(self, X, U):
        X, U = np.asarray(X, np.float64), np.asarray(U, np.float64)
        A = np.broadcast_to(self.A, X.shape[:-1] + self.A.shape[-2:])
        B = np.broadcast_to(self.B, X.shape[:-1] + self.B.shape[-2:])
        f = np.einsum("...rt,...t->...r", A, X) + np.einsum("...rt,...t->...r", B, U)
        if self.c is not None:
            f = f + self.c
        return f, A, B


class LTI(LTV):
    """Time-invariant special case ``x+ = A x + B u + c`` with ONE ``A (n,n)``, ``B (n,m)`` (optionally per particle:
    ``(M,n,n)``, ``(M,n,m)``) — broadcast over the horizon; everything else is :class:`LTV`."""

    kind = "ltv"

    def __init__(self, A, B, c=None):
        A, B = np.asarray(A, np.float64), np.asarray(B, np.float64)
        if A.ndim not in (2, 3) or B.ndim != A.ndim:
            raise ValueError("LTI expects A (n,n) / (M,n,n) and B (n,m) / (M,n,m)")
        lift = (lambda a: a[:, None]) if A.ndim == 3 else (lambda a: a)  # (M,n,n) -> (M,1,n,n): broadcasts over stages
        super().__init__(lift(A), lift(B), None if c is None else lift(np.asarray(c, np.float64)))

"""``Problem``: build the keyword arguments of ``solve`` with sensible defaults.

Behavioural counterpart of reference ``pmpc/problem_struct.py:10-155``: dimensions are inferred
from whatever arrays are given (or ``N``/``xdim``/``udim``), every array has a default, setting an
array checks its trailing shape and tiles it over the leading axes (``M`` particles), and the object
is a read-only ``Mapping`` so that ``solve(**problem)`` works.  ``Nc`` defaults to 0 and is injected
into ``solver_settings`` when ``M`` is set (``problem_struct.py:57-58,125-130``).
"""
from __future__ import annotations

from collections.abc import Mapping
from copy import copy
from typing import Dict, Tuple
from warnings import warn

import numpy as np

_DIM_MAP: Dict[str, Tuple[str, ...]] = {
    "Q": ("N", "xdim", "xdim"), "R": ("N", "udim", "udim"), "X_ref": ("N", "xdim"), "U_ref": ("N", "udim"),
    "X_prev": ("N", "xdim"), "U_prev": ("N", "udim"), "u_l": ("N", "udim"), "u_u": ("N", "udim"),
    "x_l": ("N", "xdim"), "x_u": ("N", "xdim"), "x0": ("xdim",),
}
_SCALAR_KEYS = ("solver_settings", "reg_x", "reg_u", "max_it", "res_tol", "verbose", "slew_rate", "P")
_OPTIONAL_KEYS = ("lin_cost_fn", "extra_cstrs_fns")


class Problem(Mapping):
    dim_map = _DIM_MAP

    def __init__(self, **kw):
        object.__setattr__(self, "_arrays", {})
        object.__setattr__(self, "_extra", {})
        object.__setattr__(self, "_dims", self._infer_dims(kw))
        object.__setattr__(self, "M", kw.get("M", None))
        d = self._dims
        N, xdim, udim = d["N"], d["xdim"], d["udim"]
        x0 = np.zeros(xdim)
        self._arrays.update(
            Q=np.tile(np.eye(xdim), (N, 1, 1)), R=np.tile(1e-1 * np.eye(udim), (N, 1, 1)), x0=x0,
            X_ref=np.zeros((N, xdim)), U_ref=np.zeros((N, udim)), X_prev=np.tile(x0, (N, 1)),
            U_prev=np.zeros((N, udim)), u_l=None, u_u=None, x_l=None, x_u=None)
        self._extra.update(solver_settings=dict(), reg_x=1e0, reg_u=1e0, max_it=30, res_tol=1e-6, verbose=True,
                           slew_rate=None, P=None)
        for k, v in kw.items():
            if k in ("N", "xdim", "udim", "M"):
                continue
            if k.startswith("_"):
                warn(f"Cannot set private attribute {k}")
                continue
            setattr(self, k, v)
        self._tile_all()
        if "Nc" not in self._extra:
            self._extra["Nc"] = 0

    # -- dimensions ---------------------------------------------------------------------------------
    @staticmethod
    def _infer_dims(kw):
        dims = {k: int(kw[k]) for k in ("N", "xdim", "udim") if k in kw}
        for name, axes in _DIM_MAP.items():
            if kw.get(name) is not None:
                shp = np.shape(kw[name])
                for i in range(1, len(axes) + 1):
                    dims[axes[-i]] = int(shp[-i])
        missing = [k for k in ("N", "xdim", "udim") if k not in dims]
        if missing:
            raise ValueError(f"Missing dimension {missing[0]}")
        return dims

    @property
    def dims(self):
        return copy(self._dims)

    N = property(lambda self: self._dims["N"])
    xdim = property(lambda self: self._dims["xdim"])
    udim = property(lambda self: self._dims["udim"])

    def _full_shape(self, name):
        shp = tuple(self._dims[a] for a in _DIM_MAP[name])
        return shp if self.M is None else (self.M,) + shp

    def _tile(self, name, v):
        if v is None:
            return None
        v = np.array(v)
        full = self._full_shape(name)
        assert v.shape == full[len(full) - v.ndim:], (
            f"{name} does not have the correct shape, v.shape = {v.shape}, correct_shape = {full[len(full) - v.ndim:]}")
        return np.tile(v, full[: len(full) - v.ndim] + (1,) * v.ndim)

    def _tile_all(self):
        if self.M is None:
            return
        for name in _DIM_MAP:
            self._arrays[name] = self._tile(name, self._arrays[name])
        P = self._extra.get("P")
        if P is not None:
            P = np.asarray(P)
            self._extra["P"] = P

    # -- attribute protocol -------------------------------------------------------------------------
    def __setattr__(self, k, v):
        if k in _DIM_MAP:
            self._arrays[k] = self._tile(k, v)
        elif k in ("N", "xdim", "udim"):
            raise AttributeError(f"{k} is fixed at construction")
        else:
            self._extra[k] = v

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        arrays, extra = object.__getattribute__(self, "_arrays"), object.__getattribute__(self, "_extra")
        if k in arrays:
            return arrays[k]
        if k in extra:
            return extra[k]
        raise AttributeError(k)

    def __repr__(self):
        return f"Problem({self._dims}, id={abs(hash(str(id(self))))})"

    # -- export -------------------------------------------------------------------------------------
    def to_dict(self):
        problem = {k: self._arrays[k] for k in _DIM_MAP}
        problem.update({k: self._extra.get(k) for k in _SCALAR_KEYS})
        if self.M is not None:
            ss = problem["solver_settings"]
            if "Nc" in ss and ss["Nc"] != self.Nc:
                warn(f"Nc specified in solver_settings, but Problem specifies Nc via a property. "
                     f"We will use Nc = {self.Nc} from the Problem.")
            ss["Nc"] = self.Nc
        if "f_fx_fu_fn" in self._extra:
            problem["f_fx_fu_fn"] = self._extra["f_fx_fu_fn"]
        else:
            warn("No dynamics function specified, please set `prob.f_fx_fu_fn`")
        problem.update({k: self._extra[k] for k in _OPTIONAL_KEYS if k in self._extra})
        return problem

    def __iter__(self):
        return iter(self.to_dict().keys())

    def __getitem__(self, k):
        return self.to_dict()[k]

    def __len__(self):
        return len(self.to_dict())

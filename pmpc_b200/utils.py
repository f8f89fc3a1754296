"""Small host helpers on the solve path (counterpart of reference ``pmpc/utils.py``)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


class TablePrinter:
    """ASCII iteration table with the column format of the reference's verbose output
    (``pmpc/utils.py:6-62``; used at ``scp_mpc.py:325-334,409-412``)."""

    def __init__(self, names: Sequence[str], fmts: Optional[Sequence[str]] = None, prefix: str = ""):
        self.names = list(names)
        self.fmts = list(fmts) if fmts is not None else ["%9.4e"] * len(self.names)
        self.prefix = prefix
        self.widths = [max(self._fmt_width(f), len(n)) + 2 for f, n in zip(self.fmts, self.names)]

    @staticmethod
    def _fmt_width(fmt: str) -> int:
        kind = fmt[-1]
        if kind in "fedi":
            return max(len(fmt % 1), len(fmt % -1))
        if kind == "s":
            return len(fmt % "")
        raise ValueError(f"unsupported print format {fmt!r}")

    @staticmethod
    def _center(s: str, width: int, heavy_right: bool) -> str:
        pad = width - len(s)
        assert pad >= 0
        small, big = pad // 2, pad // 2 + pad % 2
        return (" " * small + s + " " * big) if heavy_right else (" " * big + s + " " * small)

    def _rule(self) -> str:
        return "+" + "".join("-" * w + "+" for w in self.widths)

    def make_header(self) -> str:
        cells = "".join("|" + self._center(n, w, True) for n, w in zip(self.names, self.widths))
        return f"{self.prefix}{self._rule()}\n{self.prefix}{cells}|\n{self.prefix}{self._rule()}"

    def make_footer(self) -> str:
        return self.prefix + self._rule()

    def make_values(self, vals) -> str:
        assert len(vals) == len(self.fmts)
        cells = "".join("|" + self._center(f % v, w, False) for v, f, w in zip(vals, self.fmts, self.widths))
        return self.prefix + cells + "|"

    def print_header(self):
        print(self.make_header())

    def print_footer(self):
        print(self.make_footer())

    def print_values(self, vals):
        print(self.make_values(vals))


def atleast_nd(x: Optional[np.ndarray], n: int):
    """Prepend singleton axes up to ``n`` dims (``pmpc/utils.py:65-69``)."""
    if x is None:
        return None
    x = np.asarray(x)
    return x.reshape((1,) * max(n - x.ndim, 0) + x.shape)


def to_numpy_f64(x):
    if isinstance(x, (float, int)):
        return x
    return np.asarray(x, dtype=np.float64)

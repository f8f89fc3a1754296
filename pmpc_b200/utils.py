"""Host helpers of the solve path: the verbose iteration table and two array coercions.

`TablePrinter` reproduces the LOOK of the reference's verbose output (column widths, centring rules, `+---+` rules;
reference ``pmpc/utils.py:6-62``, used at ``scp_mpc.py:325-334,409-412``) so logs stay diff-able; the implementation
is table-driven: every row is rendered from per-column ``(width, bias)`` specs computed once.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np

_NUMERIC_KINDS = frozenset("fedi")


def _rendered_width(fmt: str) -> int:
    """Widest text a %-format can produce for a one-digit value (sign included)."""
    kind = fmt[-1:]
    if kind in _NUMERIC_KINDS:
        return max(len(fmt % probe) for probe in (1, -1))
    if kind == "s":
        return len(fmt % "")
    raise ValueError(f"unsupported print format {fmt!r}")


def _place(text: str, width: int, extra_on_right: bool) -> str:
    """Centre `text` in `width` columns; an odd surplus blank goes right (headers) or left (values)."""
    slack = width - len(text)
    if slack < 0:
        raise AssertionError(f"{text!r} wider than its {width}-column cell")
    left = slack // 2 if extra_on_right else slack - slack // 2
    return f"{'':{left}}{text}{'':{slack - left}}"


class TablePrinter:
    def __init__(self, names: Sequence[str], fmts: Optional[Sequence[str]] = None, prefix: str = ""):
        self.names: List[str] = [str(n) for n in names]
        self.fmts: List[str] = ["%9.4e"] * len(self.names) if fmts is None else list(fmts)
        self.prefix = prefix
        self.widths = [2 + max(_rendered_width(f), len(n)) for n, f in zip(self.names, self.fmts)]
        self._rule = self.prefix + "+" + "+".join("-" * w for w in self.widths) + "+"

    def _row(self, cells: Iterable[str], header: bool) -> str:
        body = "|".join(_place(c, w, header) for c, w in zip(cells, self.widths))
        return f"{self.prefix}|{body}|"

    def make_header(self) -> str:
        return "\n".join((self._rule, self._row(self.names, True), self._rule))

    def make_footer(self) -> str:
        return self._rule

    def make_values(self, vals) -> str:
        if len(vals) != len(self.fmts):
            raise AssertionError("one value per column expected")
        return self._row((f % v for f, v in zip(self.fmts, vals)), False)

    def print_header(self):
        print(self.make_header())

    def print_footer(self):
        print(self.make_footer())

    def print_values(self, vals):
        print(self.make_values(vals))


def atleast_nd(x: Optional[np.ndarray], n: int):
    """View of ``x`` with singleton axes prepended until it has ``n`` dimensions; None passes through."""
    if x is None:
        return None
    x = np.asarray(x)
    missing = n - x.ndim
    return x if missing <= 0 else x[(None,) * missing]


def to_numpy_f64(x):
    """Python scalars stay scalars, everything else becomes a float64 ndarray (no copy when it already is one)."""
    return x if isinstance(x, (float, int)) else np.asarray(x, dtype=np.float64)

"""Transcribes the iteration tables that the reference's OWN solver stack printed into its notebooks (PMPC.jl behind
`pmpc.solve`; outputs stored in the .ipynb files of the reference repository) into tests/golden/notebook_traces.json.

These are the only outputs of the real PMPC.jl back end that the reference repository holds for fully specified,
deterministic inputs (SURVEY.md 8c lists none; the Julia notebook PMPC.jl/test/interactive.ipynb uses unseeded randn
inputs).  Each table row is `it, elaps, obj, resid, reg_x, reg_u` printed with `%8.3e` (4 significant digits,
`pmpc/scp_mpc.py:318-321, 409-417`).  The problem each table belongs to is restated in
tests/test_reference_notebook_traces.py from the notebook cells cited here.

    python tests/golden/make_notebook_traces.py        (needs /root/reference; run in the build container)
"""
import json
import os
import re

REF = os.environ.get("PMPC_REFERENCE", "/root/reference")
SOURCES = dict(
    # name: (notebook, index of the cell whose output is the table, what the cell ran)
    gpu_solver_cpu_check=("examples/gpu_solver.ipynb", 18,
                          "pmpc.solve(**cpu_args, **cpu_opts): first problem of cells 6/11/17 - Dubins car of cell 4 (eps 1e-3, p = (1, 1, 0.3)), "
                          "N 20, Q = I, R = 1e-2 I, x0 = 1, refs / warm start 0, |u| <= 1, reg_x 3, reg_u 1, max_it 50"),
    root_testing_single=("tests/root_testing.ipynb", 4,
                         "solve(**problem) of cell 3 with p = problem['P'] = ones(N), i.e. v_scale = w_scale = T = 1 (car of cell 2, eps 1e-6): N 20, "
                         "Q = I, R = 1e-2 I, x0 = 1, |u| <= 1, reg_x = reg_u = 1, slew_rate 1e2, max_it 20, res_tol 1e-12, "
                         "solver_settings = dict(smooth_alpha=1e-1, solver='ecos')"),
    demo_cost_cpu=("tests/demo_cost_jax.ipynb", 3,
                   "first table of the cell = solve_cpu (pmpc.solve) on cells 1-2: tests/dubins_car.py model (eps 1e-6, p = (1, 1, 0.3)), "
                   "N 30, Q = I, R = 0, x0 = 1, X_ref = 0.4, warm start 0, |u| <= 1, default reg_x 1 / reg_u 1e-2, max_it 100, "
                   "solver_settings = dict(smooth_cstr='logbarrier', smooth_alpha=1e3)"),
    root_testing_consensus=("tests/root_testing.ipynb", 11,
                            "solve(**problem) of cell 10: M = 20 particles, P_i = linspace(0.7, 1, M) of shape (M, N, 1) so that p[..., 1] and "
                            "p[..., 2] clamp to p[..., 0] (JAX indexing): v_scale = w_scale = T = P_i; Nc 5, slew_rate 1e2, reg_x = reg_u = 1, "
                            "max_it 50, res_tol 1e-5, solver_settings = dict(Nc=5, solver='jump', smooth_alpha=1e0)"),
)


def table(path, cell):
    nb = json.load(open(os.path.join(REF, path)))
    text = "".join(t for o in nb["cells"][cell]["outputs"] if "text" in o for t in o["text"])
    rows = []
    for line in text.splitlines():
        if re.match(r"\| \d{4} \|", line):
            cols = [c.strip() for c in line.strip("|").split("|")]
            if int(cols[0]) != len(rows) + 1:  # a second table in the same cell (another solver): not ours
                break
            rows.append(dict(it=int(cols[0]), obj=float(cols[2]), resid=float(cols[3]), reg_x=float(cols[4]), reg_u=float(cols[5])))
    return rows


def main():
    out = {}
    for name, (path, cell, what) in SOURCES.items():
        rows = table(path, cell)
        assert rows and [r["it"] for r in rows] == list(range(1, len(rows) + 1)), name
        out[name] = dict(notebook=path, cell=cell, ran=what, rows=rows)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_traces.json")
    json.dump(out, open(dst, "w"), indent=1)
    print({k: len(v["rows"]) for k, v in out.items()}, "->", dst)


if __name__ == "__main__":
    main()

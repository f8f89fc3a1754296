"""Drop-in proof with the reference's OWN code (INTEGRATION.md level 1).

`baseline/_ref/` (built by oracle/ref_install.sh in the build container; git-ignored, ships to the GPU box) holds the
UNMODIFIED reference: its Python package `pmpc/` and its pybind11 shim `PMPC.jl/pmpcjl/module.cpp`, compiled from where
it lies and linked against `libpmpc_b200.so` in place of `libPMPC.so + libjulia.so`.  The tests run the reference's
`pmpc.solve(...)` — its `scp_solve` loop, `static_backend.aff_solve` marshalling, `pmpcjl.lqp_solve` binding — in a
subprocess with `PYTHONPATH=baseline/_ref`, so that every line above the C ABI is the reference's and everything below
it is this repository's CUDA path, and compare the result with `pmpc_b200.solve` and with the CPU oracle.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "baseline", "_ref")
rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())

RUNNER = r"""
import json, sys, numpy as np
sys.path.insert(0, {tests!r})
import pmpc                                    # the reference package (PYTHONPATH = baseline/_ref)
import pmpc.static_backend as sb
assert sb.is_precompiled_backend_available(), "reference did not find its pmpcjl extension"
assert "baseline/_ref" in pmpc.__file__.replace("\\", "/") and "baseline/_ref" in sb.pmpcjl.__file__.replace("\\", "/")
from fixtures import config_c1, config_c2
from oracle import dubins                      # NumPy Dubins model (pinned 1e-12 against the reference's torch model)
out = {{}}
for name, (d, p) in dict(c1=config_c1(), consensus=config_c2(M=6, N=12)).items():
    if name == "consensus":
        d.update(max_it=40, res_tol=1e-7)
    fn = dubins.make_f_fx_fu_fn(p)
    kw = dict(d)
    Q, R, x0 = kw.pop("Q"), kw.pop("R"), kw.pop("x0")
    X, U, data = pmpc.solve(fn, Q, R, x0, **kw)   # positional order of scp_mpc.py:205-217
    out[name] = dict(X=X.tolist(), U=U.tolist(), hist=[[h["obj"], h["resid"]] for h in data["hist"]])
json.dump(out, open(sys.argv[1], "w"))
"""


def _have_ref():
    return os.path.isdir(os.path.join(REF, "pmpc")) and any(f.startswith("pmpcjl") and f.endswith(".so") for f in os.listdir(REF))


def test_reference_shim_links_against_the_library_and_exports_its_api():
    """CPU part: the reference's compiled pmpcjl module imports (its init_julia call resolves to the library's no-op)
    and exposes lqp_solve / lcone_solve; the reference's backend probe finds it."""
    if not _have_ref():
        pytest.skip("baseline/_ref not built (oracle/ref_install.sh needs /root/reference)")
    code = ("import pmpc.static_backend as sb; assert sb.is_precompiled_backend_available(); "
            "print(sorted(n for n in dir(sb.pmpcjl) if n.endswith('_solve')))")
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=REF), capture_output=True, text=True,
                       cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "['lcone_solve', 'lqp_solve']" in r.stdout


@pytest.mark.gpu
def test_reference_solve_runs_on_the_cuda_path_and_matches(gpu, tmp_path):
    """The reference's own `pmpc.solve` (BASELINE config 1 and a small consensus case) on top of libpmpc_b200.so ==
    `pmpc_b200.solve` in host-loop mode (same C ABI calls: bit-identical) == the CPU oracle (1e-5 / 1e-6)."""
    if not _have_ref():
        pytest.fail("baseline/_ref is missing: run oracle/ref_install.sh in the build container (it ships with the snapshot)")
    lib_dir = os.path.join(REF, "PMPC.jl", "pmpcjl", "lib")
    os.makedirs(lib_dir, exist_ok=True)  # import_pmpcjl.py:12-25 lists this directory
    out = tmp_path / "ref.json"
    script = tmp_path / "run_ref.py"
    script.write_text(RUNNER.format(tests=os.path.join(ROOT, "tests")))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, ROOT]))
    r = subprocess.run([sys.executable, str(script), str(out)], env=env, capture_output=True, text=True, cwd=str(tmp_path),
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())

    import pmpc_b200
    from fixtures import config_c1, config_c2
    from oracle import dubins, scp as oscp

    for name, (d, p) in dict(c1=config_c1(), consensus=config_c2(M=6, N=12)).items():
        if name == "consensus":
            d.update(max_it=40, res_tol=1e-7)
        fn = dubins.make_f_fx_fu_fn(p)
        Xr, Ur, hr = np.array(res[name]["X"]), np.array(res[name]["U"]), np.array(res[name]["hist"])
        # (a) this repository's host loop: the same sequence of c_lqp_solve calls -> the same bits
        kw = dict(d, solver_settings=dict(d["solver_settings"], host_loop=True))
        Xh, Uh, dh = pmpc_b200.solve(fn, **kw)
        assert np.array_equal(Xr, Xh) and np.array_equal(Ur, Uh) and len(hr) == len(dh["hist"])
        # (b) the CPU oracle
        kwo = {k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "reg_x", "reg_u", "max_it", "res_tol")}
        Xo, Uo, do = oscp.scp_solve(fn, d["Q"], d["R"], d["x0"], solver_settings=d["solver_settings"], **kwo)
        assert len(hr) == len(do["hist"])
        assert rel(Xr, Xo) < 1e-5 and rel(Ur, Uo) < 1e-5
        assert abs(hr[-1, 0] - do["hist"][-1]["obj"]) <= 1e-6 * abs(do["hist"][-1]["obj"])

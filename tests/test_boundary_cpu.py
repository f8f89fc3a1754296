"""CPU tests of the drop-in boundary and host logic: the C-ABI library loads and exports every
symbol include/pmpc_b200.h declares, the product path fails loudly without a GPU (no CPU
fallback), the Python API mirrors the reference's signatures, and the N>1 sharding logic works
under a 2-process gloo group.  No compute calls are made."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pmpc_b200
from pmpc_b200 import _lib, parallel, static_backend
from pmpc_b200.utils import TablePrinter

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:void|int|long long|const char \*)\s*\*?\s*((?:pmpc|c)_[a-z0-9_]+|init_julia|shutdown_julia)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert {"c_lqp_solve", "c_lcone_solve", "pmpc_create", "pmpc_scp_solve", "pmpc_lqp_solve"} <= set(names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pmpc_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(names), set(_lib.EXPORTED_SYMBOLS) ^ set(names)
    assert b"sm_100a" in lib.pmpc_version()


def test_settings_struct_matches_the_header_defaults():
    s = _lib.default_settings()
    assert (s.rho0, s.alpha, s.max_admm_iter, s.check_every, s.polish, s.warm_start) == (1.0, 1.6, 4000, 10, 1, 1)
    assert s.eps_abs == 1e-9 and s.kkt_tol == 1e-9
    with pytest.raises(KeyError):
        _lib.default_settings(nonsense=1)


@pytest.mark.skipif(_lib.load().pmpc_device_count() > 0, reason="a GPU is present")
def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: the handle API raises, the void compat entry NaN-fills (reference failure
    convention, osqp_solver.jl:69-71 -> scp_mpc.py:391-394 returns (None, None, None))."""
    with pytest.raises(_lib.PmpcError, match="no CUDA device"):
        pmpc_b200.DeviceSolver(1, 5, 2, 1)
    M, N, n, m = 1, 4, 2, 1
    Q, R = np.tile(np.eye(n), (N, 1, 1)), np.tile(np.eye(m), (N, 1, 1))
    lti = pmpc_b200.LTV(np.tile(np.eye(n), (1, N, 1, 1)), np.ones((1, N, n, m)))
    out = pmpc_b200.solve(lti, Q, R, np.ones(n), solver_settings=dict(solver="osqp", host_loop=True), max_it=2)
    assert out == (None, None, None)
    with pytest.raises(_lib.PmpcError):
        pmpc_b200.solve(lti, Q, R, np.ones(n), solver_settings=dict(solver="osqp"), max_it=2)


def test_solve_signature_matches_reference():
    ours = list(inspect.signature(pmpc_b200.scp_solve).parameters)
    assert ours[:12] == ["f_fx_fu_fn", "Q", "R", "x0", "X_ref", "U_ref", "X_prev", "U_prev", "x_l", "x_u", "u_l", "u_u"]
    d = {k: v.default for k, v in inspect.signature(pmpc_b200.scp_solve).parameters.items()}
    assert (d["max_it"], d["res_tol"], d["reg_x"], d["reg_u"], d["slew_rate"], d["time_limit"]) == (100, 1e-5, 1.0, 1e-2, 0.0, 1000.0)
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
        import pmpc  # the reference's Python layer imports without Julia

        ref = inspect.signature(pmpc.scp_mpc.scp_solve).parameters
        assert list(ref) == ours
        for k in ref:
            if ref[k].default is not inspect._empty:
                assert ref[k].default == d[k], k
        assert pmpc.SOLVE_KWS == pmpc_b200.SOLVE_KWS
        sys.path.remove(REF)


def test_problem_defaults_and_tiling():
    p = pmpc_b200.Problem(N=5, xdim=4, udim=2, M=3, x0=np.arange(4.0), u_l=-np.ones((5, 2)))
    p.f_fx_fu_fn = pmpc_b200.DubinsCar()
    d = p.to_dict()
    assert d["Q"].shape == (3, 5, 4, 4) and d["R"].shape == (3, 5, 2, 2) and d["x0"].shape == (3, 4)
    assert d["u_l"].shape == (3, 5, 2) and d["u_u"] is None and d["solver_settings"]["Nc"] == 0
    assert np.all(d["R"][0, 0] == 0.1 * np.eye(2)) and d["reg_x"] == 1.0 and d["max_it"] == 30 and d["res_tol"] == 1e-6
    assert set(d) >= {"Q", "R", "x0", "X_ref", "U_ref", "X_prev", "U_prev", "f_fx_fu_fn", "solver_settings"}
    with pytest.raises(AssertionError):
        p.Q = np.zeros((5, 3, 3))
    with pytest.raises(ValueError):
        pmpc_b200.Problem(N=5, xdim=4)
    assert dict(**p)["reg_u"] == 1.0  # Mapping protocol: solve(**problem)


def test_table_printer_and_filters_match_reference_behaviour():
    names, fmts = ["it", "elaps", "obj", "resid", "reg_x", "reg_u"], ["%04d", "%8.3e", "%8.3e", "%8.3e", "%8.3e", "%8.3e"]
    tp = TablePrinter(names, fmts=fmts)
    row = tp.make_values((3, 0.1234, 1.5, 2e-5, 1.0, 1e-2))
    assert row.startswith("|") and row.count("|") == 7 and "0003" in row and "2.000e-05" in row
    rng = np.random.default_rng(0)
    Fs = [rng.standard_normal(12) for _ in range(4)]
    for name in ("AA", "smooth", "select"):
        w = pmpc_b200.FILTER_MAP[name](Fs)
        assert w.shape == (4,) and abs(w.sum() - 1.0) < 1e-9
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
        import pmpc
        from pmpc.utils import TablePrinter as RefTP

        rtp = RefTP(names, fmts=fmts)
        assert rtp.make_header() == tp.make_header() and rtp.make_values((3, 0.1234, 1.5, 2e-5, 1.0, 1e-2)) == row
        assert rtp.make_footer() == tp.make_footer()
        for name in ("AA", "smooth", "select"):
            assert np.allclose(pmpc.scp_mpc.FILTER_MAP[name](Fs), pmpc_b200.FILTER_MAP[name](Fs), atol=1e-9)
        sys.path.remove(REF)


def test_layout_helpers_follow_the_c_abi():
    """py2jl + asfortranarray (what the reference does, static_backend.py:71-74) and mat_cm (what
    the handle API does) give the same bytes: column-major blocks, particle slowest."""
    rng = np.random.default_rng(1)
    fx = rng.standard_normal((3, 5, 4, 4))
    a = np.asfortranarray(static_backend.py2jl(fx, 2))
    assert a.shape == (4, 4, 5, 3)
    flat_ref = a.ravel(order="K") if a.flags["F_CONTIGUOUS"] else None
    flat_ref = np.ravel(a, order="F")
    assert np.array_equal(flat_ref, _lib.mat_cm(fx).ravel())
    i, j, r, t = 2, 3, 1, 2
    assert flat_ref[r + 4 * (t + 4 * (j + 5 * i))] == fx[i, j, r, t]


def test_shard_bounds_partition():
    for total in (0, 1, 7, 16384):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from pmpc_b200 import parallel
d = parallel.init_process_group("gloo")
r, w = d.get_rank(), d.get_world_size()
lo, hi = parallel.shard_bounds(11, r, w)
parts = parallel.gather_objects(list(range(lo, hi)), d)
assert [x for p in parts for x in p] == list(range(11)), parts
assert parallel.reduce_max(float(r + 1), d) == float(w)
assert parallel.reduce_sum(float(hi - lo), d) == 11.0
d.barrier(); d.destroy_process_group()
print("ok", r)
"""


def test_two_process_gloo_sharding(tmp_path):
    """N>1 path on CPU: 2 ranks over gloo shard a batch, gather results in order, max-reduce time."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_shard_particles_copies_shared_block_bounds():
    """Particle sharding (SURVEY 8e case 2): every rank must project the shared controls with global
    particle 0's bounds (lqp_utils.jl:329-330) whatever its own slice says."""
    from pmpc_b200 import parallel

    rng = np.random.default_rng(0)
    M, N, m, Nc = 6, 5, 2, 2
    full = dict(u_l=-rng.uniform(0.5, 1.0, (M, N, m)), u_u=rng.uniform(0.5, 1.0, (M, N, m)), x0=rng.standard_normal((M, 3)),
                X_ref=None)
    parts = [parallel.shard_particles(full, *parallel.shard_bounds(M, r, 3), Nc) for r in range(3)]
    assert np.array_equal(np.concatenate([p["x0"] for p in parts]), full["x0"]) and parts[1]["X_ref"] is None
    for r, p in enumerate(parts):
        lo, hi = parallel.shard_bounds(M, r, 3)
        assert np.array_equal(p["u_l"][0, :Nc], full["u_l"][0, :Nc]) and np.array_equal(p["u_u"][0, :Nc], full["u_u"][0, :Nc])
        assert np.array_equal(p["u_l"][0, Nc:], full["u_l"][lo, Nc:]) and np.array_equal(p["u_l"][1:], full["u_l"][lo + 1:hi])


def test_outer_loop_variants_keep_the_reference_signatures():
    """accelerated_scp_solve / tune_scp are drop-ins: same parameter names, order and defaults as
    reference pmpc/accelerated.py:15-41 and pmpc/scp_mpc.py:460-467 (read from the committed signature list, the
    reference tree does not exist on the GPU box)."""
    import inspect

    import pmpc_b200

    acc = inspect.signature(pmpc_b200.accelerated_scp_solve).parameters
    assert list(acc)[:12] == ["f_fx_fu_fn", "Q", "R", "x0", "X_ref", "U_ref", "X_prev", "U_prev", "x_l", "x_u", "u_l", "u_u"]
    assert [acc[k].default for k in ("verbose", "max_it", "time_limit", "res_tol", "reg_x", "reg_u", "slew_rate", "method")] == \
        [True, 100, 1000.0, 1e-5, 1e0, 1e-2, 0.0, "lqp"]
    tune = inspect.signature(pmpc_b200.tune_scp).parameters
    assert tune["sample_nb"].default == 14 and tune["reg_rng"].default == (-3, 3) and "savefig" in tune
    assert pmpc_b200.accelerated.momentum_update(np.ones(2), np.zeros(2), 3) == pytest.approx(1.6 * np.ones(2))


def test_user_dynamics_wrapper_compiles_for_sm100a_without_a_gpu():
    """The NVRTC path of pmpc_register_dynamics, as far as it goes without a device: the wrapper kernel of
    csrc/userdyn.inl around a user function — deliberately named `f`, like the wrapper's own output array — compiles to an
    sm_100a cubin that exports `pmpc_user_linearize`; a broken source yields a log naming the user's file."""
    try:
        rt = ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc.so.12 not available")
    src = open(os.path.join(ROOT, "pmpc_b200", "csrc", "userdyn.inl")).read()
    wrapper = src[src.index('R"SRC(') + 6:src.index(')SRC";')]
    head = "#define PMPC_N 2\n#define PMPC_M 1\n#define PMPC_NX 2\n#define PMPC_NU 1\n#define PMPC_NP 1\n#define PMPC_TILE 32\n#define PMPC_USER_FN f\n#line 1 \"user_dynamics.cu\"\n"
    good = ("__device__ void f(const double* x, const double* u, const double* p, double* f, double* fx, double* fu) {"
            " f[0] = x[0] + p[0] * x[1]; f[1] = x[1] + p[0] * (u[0] - sin(x[0])); fx[0] = 1; fx[1] = p[0];"
            " fx[2] = -p[0] * cos(x[0]); fx[3] = 1; fu[1] = p[0]; }")

    def compile_(user):
        prog = ctypes.c_void_p()
        full = head + user + "\n#line 1 \"pmpc_user_wrapper.cu\"\n" + wrapper
        assert rt.nvrtcCreateProgram(ctypes.byref(prog), full.encode(), b"pmpc_user.cu", 0, None, None) == 0
        opts = (ctypes.c_char_p * 3)(b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device")
        rc = rt.nvrtcCompileProgram(prog, 3, opts)
        n = ctypes.c_size_t()
        rt.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
        log = ctypes.create_string_buffer(max(n.value, 1))
        rt.nvrtcGetProgramLog(prog, log)
        cubin = b""
        if rc == 0:
            rt.nvrtcGetCUBINSize(prog, ctypes.byref(n))
            buf = ctypes.create_string_buffer(n.value)
            rt.nvrtcGetCUBIN(prog, buf)
            cubin = buf.raw
        rt.nvrtcDestroyProgram(ctypes.byref(prog))
        return rc, log.value.decode(), cubin

    rc, log, cubin = compile_(good)
    assert rc == 0, log
    assert b"pmpc_user_linearize" in cubin and cubin[:4] == b"\x7fELF"
    rc, log, _ = compile_(good.replace("sin(x[0])", "undeclared_symbol"))
    assert rc != 0 and "user_dynamics.cu" in log and "undeclared_symbol" in log


def test_linear_rows_as_extra_states_equal_direct_rows():
    """pmpc_b200/lin_cstrs.py (host logic, no GPU): the reference's `extra_cstrs` tuples are parsed into stage-local rows,
    and the sub-problem augmented with the row values as extra states (whose box is the constraint) has the same minimiser
    as the QP with the rows taken directly into G — both solved by the CPU oracle."""
    from fixtures import rand_problem
    from oracle import lqp
    from pmpc_b200 import lin_cstrs

    rng = np.random.default_rng(30)
    M, N, n, m, L = 3, 6, 4, 2, 2
    args, kw = rand_problem(rng, M, N, n, m, True, 0.5, per_particle_slew=False)
    x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur = args
    Xf, Uf, _ = lqp.lqp_solve(*args, Nc=1, **kw)
    Gu, Gx = rng.standard_normal((M, N, L, m)), rng.standard_normal((M, N, L, n))
    val = np.einsum("ijka,ija->ijk", Gu, Uf) + np.einsum("ijkr,ijr->ijk", Gx, Xf)
    h = np.where(rng.random(val.shape) < 0.15, val - 0.1 * np.abs(val), val + 1.0)
    Xd, Ud, info = lqp.lqp_solve(*args, Nc=1, lin_rows=(Gu, Gx, h), **kw)
    assert np.sum(np.abs(np.einsum("ijka,ija->ijk", Gu, Ud) + np.einsum("ijkr,ijr->ijk", Gx, Xd) - h) < 1e-8) >= 3
    fa, fxa, fua, x0a, Xpa, Qa, Xra, xla, xua = lin_cstrs.augment((Gu, Gx, h), f, fx, fu, x0, Xp, Up, Q, R, Xr, kw["reg_x"], None, None)
    Xa, Ua, _ = lqp.lqp_solve(x0a, fa, fxa, fua, Xpa, Up, Qa, R, Xra, Ur, Nc=1, **dict(kw, lx=xla, ux=xua))
    assert np.abs(Xa[..., :n] - Xd).max() < 1e-10 and np.abs(Ua - Ud).max() < 1e-10 and np.abs(Xd - Xf).max() > 1e-2
    # reference tuple format, one particle: z = [u_1 .. u_N ; x_1 .. x_N] (README.md:232-235)
    N1, nz = 5, 5 * m + 5 * n
    G = np.zeros((2, nz))
    G[0, 2 * m:3 * m], G[0, N1 * m + 2 * n:N1 * m + 3 * n] = [1.0, -2.0], [0.5, 0.0, 0.0, 3.0]   # stage 2
    G[1, N1 * m + 4 * n + 1] = 7.0                                                                 # stage 4, state only
    Gu1, Gx1, h1 = lin_cstrs.from_extra_cstrs([(2, [], 0, G, np.zeros((2, 0)), np.array([0.3, 0.9]), np.zeros(nz), np.zeros(0))],
                                              1, N1, n, m, -1)
    assert Gu1.shape == (1, N1, 1, m) and np.array_equal(Gu1[0, 2, 0], [1.0, -2.0]) and Gx1[0, 2, 0, 3] == 3.0
    assert Gx1[0, 4, 0, 1] == 7.0 and h1[0, 2, 0] == 0.3 and h1[0, 4, 0] == 0.9 and np.isinf(h1[0, 0, 0])
    G[1, 0] = 1.0  # now row 1 couples stage 0's control with stage 4's state
    with pytest.raises(NotImplementedError, match="stage-local"):
        lin_cstrs.from_extra_cstrs([(2, [], 0, G, np.zeros((2, 0)), np.zeros(2), np.zeros(nz), np.zeros(0))], 1, N1, n, m, -1)


def test_lqp_generate_problem_matrices_equals_oracle_assembly():
    """The host-side debugging export (pmpc/scp_mpc.py:66-75 -> main.jl:374-409: P, q, A, b, G, l, u of one sub-problem)
    against the oracle's independent assembly of the same matrices: every consensus case, slew with and without the
    u_{-1} term, control and state boxes, absent bounds."""
    import scipy.sparse as sp
    from oracle.lqp import build_lqp
    from pmpc_b200 import lqp_generate_problem_matrices

    rng = np.random.default_rng(5)
    for (M, N, n, m, Nc, slew, um1, ubox, xbox) in [(3, 5, 4, 2, 2, 0.7, True, True, True), (2, 4, 3, 2, 0, 0.3, True, True, False),
                                                   (1, 6, 4, 2, 0, 0.0, False, False, True), (4, 3, 2, 1, -1, 0.5, False, True, True),
                                                   (2, 1, 3, 2, 1, 0.4, True, False, False)]:
        sym = lambda a: a @ np.swapaxes(a, -1, -2) + np.eye(a.shape[-1])
        Q, R = sym(rng.standard_normal((M, N, n, n))), sym(rng.standard_normal((M, N, m, m)))
        f, fx, fu = rng.standard_normal((M, N, n)), rng.standard_normal((M, N, n, n)), rng.standard_normal((M, N, n, m))
        Xp, Up, Xr, Ur = (rng.standard_normal((M, N, d)) for d in (n, m, n, m))
        x0 = rng.standard_normal((M, n))
        kw = dict(reg_x=rng.uniform(0.5, 1.5, M), reg_u=rng.uniform(0.01, 0.1, M), Nc=Nc)
        if slew:
            kw["slew_reg"] = np.full(M, slew)
        if um1:
            kw.update(slew_reg0=rng.uniform(0.1, 0.5, M), slew_um1=rng.standard_normal((M, m)))
        if ubox:
            kw.update(lu=-rng.uniform(0.5, 1, (M, N, m)), uu=rng.uniform(0.5, 1, (M, N, m)))
        if xbox:
            kw.update(lx=-rng.uniform(2, 3, (M, N, n)), ux=rng.uniform(2, 3, (M, N, n)))
        o = build_lqp(x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur, **kw)
        P, q, A, b, G, l, u = lqp_generate_problem_matrices(x0, f, fx, fu, Xp, Up, Q, R, Xr, Ur, **kw)
        for mine, ref in ((P, o.P), (A, o.A), (G, o.G)):
            assert mine.shape == ref.shape and abs(sp.csc_matrix(mine) - ref).max() < 1e-13 if mine.shape[0] else mine.shape == ref.shape
        assert np.abs(q - o.q).max() < 1e-13 and np.abs(b - o.b).max() < 1e-13
        assert np.array_equal(l, o.l) and np.array_equal(u, o.u)
    # the reference wrapper's call shape: no settings, a single particle without the leading axis
    P, q, A, b, G, l, u = lqp_generate_problem_matrices(x0[0], f[0], fx[0], fu[0], Xp[0], Up[0], Q[0], R[0], Xr[0], Ur[0])
    assert P.shape == (N * (n + m), N * (n + m)) and A.shape == (N * n, N * (n + m)) and G.shape[0] == 0


def test_batch_grouping_signature_separates_what_must_not_share_a_launch():
    """`solve_problems` groups problem dictionaries by `batch._sig`; everything the group's single launch decides once has
    to be part of it (an LTV model with and one without an affine term, device-solver settings, cone / smoothing requests,
    shapes, stop rule), and whatever `solve` itself routes to its host loop must not be grouped at all (-> None)."""
    from pmpc_b200 import LTV, DubinsCar
    from pmpc_b200.batch import _sig

    M, N, n, m = 2, 5, 4, 2
    A, Bm, c = np.tile(np.eye(n), (M, N, 1, 1)), np.zeros((M, N, n, m)), np.ones((M, N, n))
    base = dict(Q=np.tile(np.eye(n), (M, N, 1, 1)), R=np.tile(np.eye(m), (M, N, 1, 1)), x0=np.ones((M, n)),
                u_l=-np.ones((M, N, m)), u_u=np.ones((M, N, m)), solver_settings=dict(solver="osqp", Nc=1))
    lin, aff = dict(base, f_fx_fu_fn=LTV(A, Bm)), dict(base, f_fx_fu_fn=LTV(A, Bm, c))
    assert _sig(lin) is not None and _sig(lin) == _sig(dict(base, f_fx_fu_fn=LTV(2 * A, Bm)))  # same structure, other numbers
    assert _sig(lin) != _sig(aff)
    assert _sig(lin) != _sig(dict(lin, solver_settings=dict(solver="osqp", Nc=1, b200=dict(fused=0))))
    assert _sig(lin) != _sig(dict(lin, solver_settings=dict(solver="ecos", Nc=1)))
    assert _sig(lin) != _sig(dict(lin, solver_settings=dict(solver="osqp", Nc=1, smooth_alpha=10.0)))
    assert _sig(lin) != _sig(dict(lin, solver_settings=dict(solver="osqp", Nc=2)))
    assert _sig(lin) != _sig(dict(lin, max_it=7)) and _sig(lin) != _sig(dict(lin, res_tol=1e-3))
    assert _sig(lin) != _sig(dict(lin, x_l=-np.ones((M, N, n)), x_u=np.ones((M, N, n))))
    assert _sig(lin) != _sig(dict(lin, reg_x=2.0)) and _sig(lin) != _sig(dict(lin, slew_rate=0.5))
    car = dict(base, f_fx_fu_fn=DubinsCar(np.array([[1.0, 1.0, 0.3]] * M)))
    assert _sig(car) is not None and _sig(car) != _sig(lin)
    for host_only in (dict(solver_settings=dict(solver="osqp", Nc=1, host_loop=True)), dict(debug=True), dict(return_min_viol=True),
                      dict(time_limit=5.0), dict(filter_method="AA"), dict(lin_cost_fn=lambda X, U, problems: (0 * X, 0 * U)),
                      dict(extra_cstrs_fns=[lambda X, U, problems: []]), dict(f_fx_fu_fn=lambda X, U: (X, None, None))):
        assert _sig(dict(lin, **host_only)) is None, host_only


def test_extra_cstrs_tuples_are_mapped_to_stage_rows_or_refused_by_name():
    """`lin_cstrs.from_extra_cstrs` reads the reference's 8-tuples `(l, q, e, G_left, G_right, h, c_left, c_right)`
    (README.md:219-239 of the reference, variable vector z = [controls; states]): stage-local linear rows become
    (Gu, Gx, h) per (particle, stage); cones, new variables, cost augmentation and rows that couple stages are refused."""
    from pmpc_b200.lin_cstrs import from_extra_cstrs

    M, N, n, m, Nc = 2, 3, 2, 1, 1
    Nf = N - Nc
    nu_tot = Nc * m + M * Nf * m
    nz = nu_tot + M * N * n
    ucol = lambda i, j: j * m if j < Nc else Nc * m + i * Nf * m + (j - Nc) * m
    xcol = lambda i, j: nu_tot + i * N * n + j * n
    G = np.zeros((2, nz))
    G[0, ucol(1, 2)], G[0, xcol(1, 2) + 1] = 0.5, -2.0   # particle 1, stage 2: 0.5 u - 2 x_1 <= 3
    G[1, xcol(0, 0)] = 1.0                               # particle 0, stage 0: x_0 <= 4
    tup = (2, [], 0, G, np.zeros((2, 0)), np.array([3.0, 4.0]), np.zeros(nz), np.zeros(0))
    Gu, Gx, h = from_extra_cstrs([tup], M, N, n, m, Nc)
    assert Gu.shape == (M, N, 1, m) and Gx.shape == (M, N, 1, n) and h.shape == (M, N, 1)
    assert Gu[1, 2, 0, 0] == 0.5 and Gx[1, 2, 0, 1] == -2.0 and h[1, 2, 0] == 3.0
    assert Gx[0, 0, 0, 0] == 1.0 and h[0, 0, 0] == 4.0
    assert np.isinf(h).sum() == M * N - 2 and np.count_nonzero(Gu) == 1 and np.count_nonzero(Gx) == 2
    bad = {
        "cone": (0, [3], 0, np.zeros((3, nz)), np.zeros((3, 0)), np.zeros(3), np.zeros(nz), np.zeros(0)),
        "new variables": (1, [], 0, G[:1], np.ones((1, 2)), np.ones(1), np.zeros(nz), np.zeros(0)),
        "cost augmentation": (1, [], 0, G[:1], np.zeros((1, 0)), np.ones(1), np.ones(nz), np.zeros(0)),
    }
    for what, t in bad.items():
        with pytest.raises(NotImplementedError, match=what.split()[0]):
            from_extra_cstrs([t], M, N, n, m, Nc)
    G2 = np.zeros((1, nz))
    G2[0, xcol(0, 0)], G2[0, xcol(0, 1)] = 1.0, 1.0  # couples two stages
    with pytest.raises(NotImplementedError, match="stage"):
        from_extra_cstrs([(1, [], 0, G2, np.zeros((1, 0)), np.ones(1), np.zeros(nz), np.zeros(0))], M, N, n, m, Nc)
    G3 = np.zeros((1, nz))
    G3[0, ucol(0, 0)] = 1.0  # a row on the shared control of a multi-particle problem
    with pytest.raises(NotImplementedError, match="shared"):
        from_extra_cstrs([(1, [], 0, G3, np.zeros((1, 0)), np.ones(1), np.zeros(nz), np.zeros(0))], M, N, n, m, Nc)


def test_batched_result_records_follow_the_reference_format(monkeypatch):
    """batch.solve_same_shape (host logic, a stand-in for the device handle): one (X, U, data) per problem with X including x0,
    `data["hist"]` one dict per SCP iteration of THAT problem with the reference's keys (scp_mpc.py:409-417), per-problem
    device counters, (None, None, None) for a problem the device reports as failed, and the single-particle shapes."""
    from pmpc_b200 import batch

    its = np.array([3, 5, 2])
    seen = {}

    class Handle:
        def __init__(self, M, N, n, m, Nc, B=1, device=-1):
            self.B, self.M, self.N, self.n, self.m, self.Nc = B, M, N, n, m, Nc

        def set_settings(self, **kw):
            seen["settings"] = kw

        def set_problem(self, x0, Q, R, *rest):
            seen["x0"], seen["Q"] = np.array(x0), np.array(Q)

        def set_warm_start(self, Xp, Up):
            seen["warm"] = (Xp, Up)

        def set_ltv(self, A, B, c):
            seen["ltv"] = (A.shape, B.shape, c)

        def scp_solve(self, max_it, res_tol):
            seen["stop"] = (max_it, res_tol)

        def solution(self):
            X = np.arange(self.B * self.M * self.N * self.n, dtype=float).reshape(self.B, self.M, self.N, self.n)
            return X, -np.ones((self.B, self.M, self.N, self.m))

        def info(self):
            return dict(status=np.array([0, 0, 1]), scp_iters=its, factorizations=np.array([7, 8, 9]))

        def hist(self, k):
            assert k == 5
            t = np.arange(5)[:, None] + 10.0 * np.arange(3)[None, :]
            return dict(obj=t, resid=-t, admm_iters=(2 * t).astype(np.int64), kkt=t)

    monkeypatch.setattr(batch, "DeviceSolver", Handle)
    N, n, m = 4, 3, 2
    for M in (None, 2):
        lead = () if M is None else (M,)
        probs = [dict(f_fx_fu_fn=pmpc_b200.LTI(np.eye(n), np.ones((n, m))), Q=np.tile(np.eye(n), lead + (N, 1, 1)),
                      R=np.tile(np.eye(m), lead + (N, 1, 1)), x0=np.full(lead + (n,), 100.0 + b), reg_x=2.0, reg_u=0.5, max_it=9,
                      res_tol=1e-3, solver_settings=dict(Nc=1, solver="osqp")) for b in range(3)]
        out, ds = batch.solve_same_shape(probs)
        assert isinstance(ds, Handle) and ds.B == 3 and seen["stop"] == (9, 1e-3) and seen["warm"] == (None, None)
        assert out[2] == (None, None, None)
        for b in range(2):
            X, U, data = out[b]
            assert X.shape == lead + (N + 1, n) and U.shape == lead + (N, m)
            assert np.all(X[..., 0, :] == 100.0 + b) and np.all(U == -1.0)
            full = np.arange(3 * (M or 1) * N * n, dtype=float).reshape(3, M or 1, N, n)[b]
            assert np.array_equal(X[..., 1:, :], full if M else full[0])
            assert [h["it"] for h in data["hist"]] == list(range(1, its[b] + 1))
            assert [h["obj"] for h in data["hist"]] == [i + 10.0 * b for i in range(its[b])]
            assert [h["resid"] for h in data["hist"]] == [-(i + 10.0 * b) for i in range(its[b])]
            assert all(set(h) == {"it", "elaps", "obj", "resid", "reg_x", "reg_u"} and (h["reg_x"], h["reg_u"]) == (2.0, 0.5)
                       for h in data["hist"])
            assert [s["admm_iters"] for s in data["solver_data"]] == [int(2 * (i + 10 * b)) for i in range(its[b])]
            assert len(data["t_aff_solve"]) == its[b] and data["device"] == dict(status=0, scp_iters=int(its[b]), factorizations=7 + b)
            assert all(type(h["obj"]) is float and type(h["it"]) is int for h in data["hist"])

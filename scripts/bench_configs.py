"""Times every BASELINE.json config at full size on one GPU through the public API (host arrays in, host
arrays out) and prints one JSON line per config.  bench.py stays the contract benchmark (C2); this is the
side table DESIGN.md quotes for C1, C3, C4, C5.     usage: python scripts/bench_configs.py [c1 c3 c4 c4lin c5]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import pmpc_b200
from pmpc_b200 import DubinsCar, LTV
from fixtures import config_c1, config_c3, config_c4, config_c5

which = sys.argv[1:] or ["c1", "c3", "c4", "c4lin", "c5"]


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)), out


if "c1" in which:
    d, p = config_c1()
    kw = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    car = DubinsCar(p[0])
    t, (X, U, data) = timed(lambda: pmpc_b200.solve(car, **kw), 50)  # plain calls: handles are cached inside solve
    print(json.dumps(dict(config="C1 Dubins N=20 M=1, 10 SCP iterations", p50_ms=1e3 * t, kernel_ms=data["device"]["kernel_ms"],
                          scp_iters=len(data["hist"]), solves_per_s=1.0 / t)), flush=True)

if "c3" in which:
    B = int(os.environ.get("C3_B", 16384))
    d, p = config_c3(B=B)
    args = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                max_it=d["max_it"], res_tol=d["res_tol"], reg_x=d["reg_x"], reg_u=d["reg_u"], slew_rate=0.0, Nc=1)
    Q, R = np.ascontiguousarray(d["Q"]), np.ascontiguousarray(d["R"])
    st = {}
    def run():
        X, U, info, ds = pmpc_b200.solve_batch(p, Q, R, d["x0"], solver=st.get("ds"), reuse_buffers=True,
                                               chunks=int(os.environ.get("C3_CHUNKS", 6)), **args)
        st["ds"] = ds
        return X, U, info, ds
    t, (X, U, info, ds) = timed(run, 3)
    print(json.dumps(dict(config="C3 %d independent Dubins problems N=30 M=16 Nc=1" % B, e2e_s=t, solves_per_s=B / t,
                          kernel_ms=ds.last_kernel_ms(), device_solves_per_s=B / (ds.last_kernel_ms() * 1e-3),
                          chunks=int(os.environ.get("C3_CHUNKS", 6)),
                          scp_iters=[int(info["scp_iters"].min()), float(info["scp_iters"].mean()), int(info["scp_iters"].max())],
                          status_ok=bool((info["status"] == 0).all()), consensus=bool(np.all(U[:, :, 0] == U[:, :1, 0])))), flush=True)

def fp64_peak():
    import ctypes as C
    from pmpc_b200 import _lib
    r = C.c_double()
    return r.value if _lib.load().pmpc_microbench(0, 1, C.byref(r)) == 0 else float("nan")


for name, cfg, label in (("c4", config_c4, "C4 LTV n=12 m=4 N=100 M=64 Nc=1, slew, u and x boxes"),
                         ("c5", config_c5, "C5 LTV n=32 m=8 N=200 M=1024 Nc=5, slew, u box (one GPU)")):
    if name not in which:
        continue
    d, A, Bm = cfg()
    fn = LTV(A, Bm)
    t, (X, U, data) = timed(lambda: pmpc_b200.solve(fn, **d), 3 if name == "c4" else 2)
    dev = data["device"]
    M, N = U.shape[:2]
    n, m = A.shape[-1], Bm.shape[-1]
    Nc = d["solver_settings"]["Nc"]
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    # flops = M N [F (4n^3 + 6n^2 m + 4n m^2 + m^3/3) + I (6n^2 + 8nm + 2m^2)], F = I = active-set passes here
    flops = M * N * dev["polish_passes"] * ((4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3.0) + (6 * n * n + 8 * n * m + 2 * m * m))
    peak = fp64_peak()
    print(json.dumps(dict(config=label, p50_s=t, kernel_ms=dev["kernel_ms"], scp_iters=len(data["hist"]), resid=data["hist"][-1]["resid"],
                          status=dev["status"], admm_iters=dev["admm_iters"], factorizations=dev["factorizations"], polish=dev["polish_passes"],
                          dyn_violation=float(np.abs(f - X[:, 1:]).max()), consensus=bool(np.all(U[:, :Nc] == U[:1, :Nc])),
                          u_in_box=bool(np.abs(U).max() <= 1 + 1e-9), kkt=max(sd["kkt"] for sd in data["solver_data"]),
                          roofline=dict(bound="fp64", flops_survey_8d=flops, achieved_tflops=flops / (dev["kernel_ms"] * 1e-3) / 1e12,
                                        peak_tflops_measured_here=peak, frac=flops / (dev["kernel_ms"] * 1e-3) / 1e12 / peak))), flush=True)


if "c4lin" in which:
    # BASELINE config 4 exactly as specified: the state box AND n_lin = 8 random linear rows g'[u; x] <= h per (particle, stage).
    # The rows travel as 8 extra states whose upper bound is h (pmpc_b200/lin_cstrs.py, host SCP loop: one device QP per
    # iteration, 20-state shape); the device solver carries the active ones by the multiplier method inside the
    # warp-per-particle kernel (wpp.cuh xbox_terms).
    d, A, Bm = config_c4(n_lin=8)
    fn = LTV(A, Bm)
    t, (X, U, data) = timed(lambda: pmpc_b200.solve(fn, **d), 3)
    Gu, Gx, h = d["solver_settings"]["lin_cstrs"]
    val = np.einsum("ijka,ija->ijk", Gu, U) + np.einsum("ijkr,ijr->ijk", Gx, X[:, 1:])
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    print(json.dumps(dict(config="C4 LTV n=12 m=4 N=100 M=64 Nc=1, slew, u and x boxes, n_lin: 8", n_lin=8, p50_s=t, scp_iters=len(data["hist"]),
                          resid=data["hist"][-1]["resid"], device_qp_s_mean=float(np.mean(data["t_aff_solve"])),
                          row_violation=float((val - h).max()), active_rows=int(np.sum(val >= h - 1e-7)), rows=int(h.size),
                          dyn_violation=float(np.abs(f - X[:, 1:]).max()), consensus=bool(np.all(U[:, :1] == U[:1, :1])),
                          u_in_box=bool(np.abs(U).max() <= 1 + 1e-9))), flush=True)

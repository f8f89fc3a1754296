#!/usr/bin/env python
"""Benchmark of the particle-SCP solve path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPUs

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): Dubins car, xdim=4, udim=2, N=50, M=128
perturbed-dynamics particles, Nc=1 shared control, box control limits, res_tol=1e-5, max_it=100.
One *step* = one batch of B independent C2 problems (problem 0 is exactly the seed-1 fixture, the
others perturb x0 and the particle parameters) solved to the stop rule; the unit is "solves".
Under torchrun every rank owns its own batch of B problems (weak scaling, no data-path collective:
SURVEY.md 8e case 1).

`value`  : solves/s with the inputs already resident in HBM (per step: device-side warm-start reset
           + the SCP kernel), timed with CUDA events on the solver's stream, max over ranks.
`e2e`    : the same batch through the public API `pmpc_b200.solve_batch` with HOST arrays: H2D of
           every input, the solve, D2H of X, U and the per-problem counters, wall clock, max over ranks.
`roofline`: the SCP kernel against the measured HBM bandwidth, algorithmic bytes per DESIGN.md.
`cpu_baseline`: the reference algorithm restated in plain C (oracle/c: lqp_utils.jl assembly + OSQP 0.6.2 at the
           reference's default settings), one problem per host thread on all cores, bounded sample —
           reported beside, not the target.
Only the cpu_baseline / --impl reference legs touch `oracle/`; the product path has no CPU fallback.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, N, NX, NU, NC = 128, 50, 4, 2, 1
MAX_IT, RES_TOL = 100, 1e-5
METRIC = "particle-SCP solves/sec (Dubins N=50 M=128 Nc=1, box control limits, res_tol 1e-5)"
WORKLOAD = "C2 Dubins N=50 M=128 Nc=1 particle consensus"
WORKLOADS = {  # BASELINE.json configs[1] (the one `metric` is quoted on) and configs[2]
    "c2": dict(M=128, N=50, metric=METRIC, name=WORKLOAD),
    "c3": dict(M=16, N=30, metric="particle-SCP solves/sec (batched independent MPC: 16384 Dubins problems N=30 M=16 Nc=1, box control "
                                  "limits, res_tol 1e-5)", name="C3 16384 independent Dubins problems N=30 M=16 Nc=1, sharded over the GPUs"),
}


def set_workload(name):
    global M, N, METRIC, WORKLOAD
    w = WORKLOADS[name]
    M, N, METRIC, WORKLOAD = w["M"], w["N"], w["metric"], w["name"]


def config_dict():
    """Identical in both arms (the driver compares it): the workload and its stop rule, and how L2 is kept cold."""
    return {"workload": WORKLOAD, "max_it": MAX_IT, "res_tol": RES_TOL,
            "l2": "inputs larger than L2: the batch working set is several GB per GPU (> 126 MB L2); no flush needed"}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def c2_iters_fixture():
    with open(os.path.join(ROOT, "tests", "golden", "c2_reference_iters.json")) as fh:
        return json.load(fh)


def reference_table_check():
    """Outside the timed region: the one problem for which the reference repository holds output of its REAL solver stack
    with the hard box the kernels solve (the table `pmpc.solve` printed in examples/gpu_solver.ipynb cell 18, transcribed
    into tests/golden/notebook_traces.json), solved here through the public call and compared at the printed precision
    (same assertions as tests/test_reference_notebook_traces.py).  Never raises: a failure is reported in the line."""
    try:
        import pmpc_b200

        with open(os.path.join(ROOT, "tests", "golden", "notebook_traces.json")) as fh:
            rows = json.load(fh)["gpu_solver_cpu_check"]["rows"]
        n_, m_, N_ = 4, 2, 20
        X, U, data = pmpc_b200.solve(pmpc_b200.DubinsCar(p=np.array([1.0, 1.0, 0.3]), eps=1e-3), np.tile(np.eye(n_), (N_, 1, 1)),
                                     np.tile(1e-2 * np.eye(m_), (N_, 1, 1)), np.ones(n_), np.zeros((N_, n_)), np.zeros((N_, m_)),
                                     np.zeros((N_, n_)), np.zeros((N_, m_)), u_l=-np.ones((N_, m_)), u_u=np.ones((N_, m_)),
                                     reg_x=3.0, reg_u=1.0, max_it=50, res_tol=1e-5)
        if data is None or len(data["hist"]) != len(rows):
            return {"ok": False, "rows": len(rows), "rows_here": None if data is None else len(data["hist"])}
        obj = np.array([h["obj"] for h in data["hist"]]) / np.array([r["obj"] for r in rows]) - 1.0
        res = np.array([h["resid"] for h in data["hist"]]) / np.array([r["resid"] for r in rows]) - 1.0
        out = {"rows": len(rows), "max_rel_obj": float(np.abs(obj).max()), "max_rel_resid_rows_1_12": float(np.abs(res[:12]).max()),
               "max_rel_resid_later_rows_but_30_31": float(np.abs(np.delete(res, [29, 30])).max()), "printed_digits": 4,
               "source": "reference examples/gpu_solver.ipynb cell 18 (PMPC.jl output)"}
        out["ok"] = bool(out["max_rel_obj"] <= 3.5e-4 and out["max_rel_resid_rows_1_12"] <= 5e-4
                         and out["max_rel_resid_later_rows_but_30_31"] <= 2.5e-2)
        return out
    except Exception as exc:  # the bench line must still be printed
        return {"ok": None, "error": "%s: %s" % (type(exc).__name__, exc)}


def make_batch(B: int, seed: int):
    """B C2-shaped problems as host arrays (Python layout of pmpc.solve with a leading B axis)."""
    from fixtures import config_c2

    d0, p0 = config_c2(M, N)
    rng = np.random.default_rng(1000 + seed)
    p = np.stack([rng.uniform(0.7, 1.0, (B, M)), rng.uniform(0.7, 1.0, (B, M)), np.full((B, M), 0.3)], -1)
    x0 = np.repeat(1.0 + 0.05 * rng.standard_normal((B, 1, NX)), M, axis=1)
    if seed == 0:
        p[0], x0[0] = p0, d0["x0"]
    tile = lambda a: np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape))
    arrs = dict(Q=tile(d0["Q"]), R=tile(d0["R"]), x0=np.ascontiguousarray(x0), X_ref=tile(d0["X_ref"]),
                U_ref=tile(d0["U_ref"]), X_prev=tile(d0["X_prev"]), U_prev=tile(d0["U_prev"]), u_l=tile(d0["u_l"]),
                u_u=tile(d0["u_u"]))
    if M == 16:  # C3: X_ref = 0.1 N(0,1) per problem, x0 ~ N(1, 0.1^2)
        arrs["X_ref"] = np.ascontiguousarray(np.repeat(0.1 * rng.standard_normal((B, 1, N, NX)), M, axis=1))
        arrs["x0"] = np.ascontiguousarray(np.repeat(1.0 + 0.1 * rng.standard_normal((B, 1, NX)), M, axis=1))
    return arrs, np.ascontiguousarray(p), d0


def untiled(arrs):
    """The same batch as `make_batch` builds, with every input that is identical for all particles given once: one block
    (N, .) for the whole batch (costs, bounds, warm start), one block per problem for X_ref, one x0 per problem.  This is
    what the end-to-end leg uploads (solve_batch expands on the device, pmpc_set_broadcast)."""
    out = {}
    for k, a in arrs.items():
        if k == "x0":
            out[k] = np.ascontiguousarray(a[:, 0])
        elif k == "X_ref" and M == 16:
            out[k] = np.ascontiguousarray(a[:, 0])
        else:
            assert np.array_equal(a[0, 0], a[-1, -1])
            out[k] = np.ascontiguousarray(a[0, 0])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# algorithmic work per (particle, stage), DESIGN.md section 5
def alg_bytes_flops(info, n=NX, m=NU, nlin=10):
    """HBM bytes the fused two-sweep pass has to move per (particle, stage): the backward sweep reads X_prev, U_prev
    and the active set and writes the hand-off [K | k | H | sparse stage]; the forward sweep reads X_prev, U_prev, the
    hand-off and the active set and writes X, U and the active set.  Everything else (costs, bounds, P, S, G, the
    Dubins linearisation) stays in registers / shared memory.  486 B at n=4, m=2 (10 of the 14 sparse-stage entries and
    the off-diagonal of H are handed over; the rest is rebuilt from the particle's parameters / folded into k).  Flops: the dense Riccati count of
    SURVEY.md 8(d) (the sparse Dubins stage needs fewer; the figure is an upper bound on useful work)."""
    passes = info["polish_passes"].astype(np.float64)
    I = (info["admm_iters"] + info["polish_passes"]).astype(np.float64)
    handoff = n * m + m + m * (m - 1) // 2 + nlin  # K, k, off-diagonals of H, sparse stage
    b_pass = 8 * (3 * (n + m) + 2 * handoff) + 3 * m
    f_fac = 4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3.0
    f_vec = 6 * n * n + 8 * n * m + 2 * m * m
    per = M * N
    return float(per * (passes * b_pass).sum()), float(per * (passes * f_fac + I * f_vec).sum())


def survey_resident_bytes(info, n=NX, m=NU):
    """SURVEY.md 8(d)'s figure for a problem whose factors stay on chip: the problem data
    8 (2 n^2 + n m + m^2 + 4 n + 4 m) = 544 B per (particle, stage) streamed ONCE per SCP iteration."""
    return float(M * N * 8 * (2 * n * n + n * m + m * m + 4 * n + 4 * m) * info["scp_iters"].astype(np.float64).sum())


def survey_streaming_bytes(info, n=NX, m=NU, nparam=3):
    """SURVEY.md 8(d)'s *streaming* model (every factorisation streams Q, fx, fu, R in and K, P out; every vector pass
    streams fx, fu, K and the stage vectors; every SCP iteration streams the linearisation once) — what an unfused
    implementation would move.  Reported beside the roofline as the traffic the fusion removed."""
    F = info["factorizations"].astype(np.float64)
    I = (info["admm_iters"] + info["polish_passes"]).astype(np.float64)
    S = info["scp_iters"].astype(np.float64)
    b_fac = 8 * (3 * n * n + 2 * n * m + m * m)
    b_vec = 8 * (n * n + 2 * n * m + m * m) + 16 * (n + m)
    b_lin = 8 * (n + m + nparam) + 8 * (n + n * n + n * m)
    return float(M * N * (F * b_fac + I * b_vec + S * b_lin).sum())


# ----------------------------------------------------------------------------------------------------
def cpu_sample(cores: int, its: int, seed: int = 0, whole: bool = False):
    """CPU leg: `cores` problems of the workload, one per host thread, each the first `its` SCP iterations (res_tol = -1
    forces exactly `its`) — or, with `whole`, each solved to the workload's stop rule — with the reference's algorithm
    restated in C (oracle/c/pmpc_oracle.c: lqp_utils.jl assembly + OSQP 0.6.2 at its defaults + Dubins + the SCP loop).
    Returns wall seconds (and the iteration counts with `whole`)."""
    from oracle import cport

    arrs, p, d0 = make_batch(cores, seed)
    t = time.perf_counter()
    X, U, iters = cport.scp_dubins_batch(cores, p, arrs["Q"], arrs["R"], arrs["x0"], arrs["X_ref"], arrs["U_ref"], arrs["X_prev"],
                                         arrs["U_prev"], arrs["u_l"], arrs["u_u"], reg_x=d0["reg_x"], reg_u=d0["reg_u"],
                                         max_it=MAX_IT if whole else its, res_tol=RES_TOL if whole else -1.0, Nc=NC)
    sec = time.perf_counter() - t
    assert np.isfinite(X).all() and (whole or (iters == its).all())
    return (sec, np.asarray(iters)) if whole else sec


def cpu_whole_solves(cores: int, rounds: int, seed: int = 0):
    """`rounds` x `cores` whole solves (C3: a problem is 480 particle-stages, a few seconds per core).  Returns
    (solves/s, sample description)."""
    sec, its = 0.0, []
    for r in range(rounds):
        s, i = cpu_sample(cores, 0, seed=seed + 100 * r, whole=True)
        sec += s
        its.append(i)
    its = np.concatenate(its)
    return cores * rounds / sec, (f"{cores} host threads x {rounds} WHOLE solves each of problems of this workload (res_tol {RES_TOL:g}, "
                                  f"max_it {MAX_IT}; {its.mean():.1f} SCP iterations on average) with the reference's solver restated "
                                  f"in C (oracle/c/pmpc_oracle.c, OSQP 0.6.2 defaults), {sec:.1f} s wall")


def reference_arm(args):
    """The reference's CPU algorithm for the path on all host cores (the Julia/OSQP stack itself is not buildable in
    this image, DESIGN.md: the plain-C restatement under oracle/c stands in, `kind: "port"`).  Each step: every host
    thread runs the first `sample_its` SCP iterations of its own C2 problem; a full solve takes `full_its` iterations
    (tests/golden/c2_reference_iters.json, measured with the same code), so
    solves/s = cores * (sample_its / full_its) / step_time."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_sample(cores, 1)  # warm-up: caches and thread pool only; one SCP iteration each
    if args.workload == "c3":  # small problems: every step is one whole solve per host thread
        t0 = time.perf_counter()
        value, sample = cpu_whole_solves(cores, args.steps, seed=1)
        step = (time.perf_counter() - t0) / args.steps
    else:
        full_its = c2_iters_fixture()["c_port_scp_iters"]
        sample_its = args.ref_sample_its
        times = [cpu_sample(cores, sample_its, seed=k) for k in range(args.steps)]
        step = float(np.mean(times))
        value = cores * (sample_its / full_its) / step
        sample = (f"{cores} host threads, each the first {sample_its} of {full_its} SCP iterations of its own C2 problem "
                  f"(plain-C restatement: OSQP 0.6.2 at the reference's defaults eps=1e-3), scaled to full solves")
    import shutil

    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(),
            "details": {"host_cores": cores, "julia_on_this_box": shutil.which("julia"),
                        "reference": "restated CPU baseline (oracle/c), not PMPC.jl: Julia/OSQP absent from the image"},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="independent C2 problems per GPU per step (0 = --waves full waves: SMs x resident CTAs per SM x waves)")
    ap.add_argument("--waves", type=int, default=8)
    ap.add_argument("--latency-solves", type=int, default=200)
    ap.add_argument("--e2e-chunks", type=int, default=0, help="sub-batches of the pipelined end-to-end leg (1 = no pipelining; 0 = one sub-batch per full wave of the solve kernel, at most 16: each costs a handle's worth of host work)")
    ap.add_argument("--cpu-sample-its", type=int, default=16)
    ap.add_argument("--ref-sample-its", type=int, default=16,
                    help="SCP iterations every host thread runs per step in the reference arm (62 = whole solves; early iterations "
                         "are the cheaper ones, so a short sample favours the CPU)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE configs[1] (the metric's config), c3 = configs[2]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == "reference":
        return reference_arm(args)

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import pmpc_b200
    from pmpc_b200 import DeviceSolver, _lib, parallel

    lib = _lib.load()
    if lib.pmpc_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — pmpc_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch

        torch.cuda.set_device(local)
        dist = parallel.init_process_group("nccl")
    dev_t = None
    if dist is not None:
        import torch

        dev_t = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 0)
    B = args.batch
    if B <= 0 and args.workload == "c3":
        B = 16384 // world  # BASELINE configs[2]: 16384 problems sharded over the GPUs (strong scaling)
    if B <= 0:
        # one CTA per problem: size the batch in whole waves of the solve kernel (148 SMs x resident CTAs)
        a1, p1, d1 = make_batch(1, 0)
        with DeviceSolver(M, N, NX, NU, NC, B=1, device=local) as probe:
            probe.set_problem(a1["x0"], a1["Q"], a1["R"], a1["X_ref"], a1["U_ref"], None, None, a1["u_l"], a1["u_u"],
                              d1["reg_x"], d1["reg_u"], np.zeros(M))
            probe.set_dubins(p1)
            ctas, sms = probe.occupancy()
        B = ctas * sms * args.waves
    arrs, p, d0 = make_batch(B, rank)
    # the e2e leg copies from page-locked host memory (the driver contract's "pinned host memory")
    arrs = {k: pmpc_b200.pinned_copy(v) for k, v in arrs.items()}
    p = pmpc_b200.pinned_copy(p)
    ds = DeviceSolver(M, N, NX, NU, NC, B=B, device=local)
    ds.set_problem(arrs["x0"], arrs["Q"], arrs["R"], arrs["X_ref"], arrs["U_ref"], None, None, arrs["u_l"], arrs["u_u"],
                   d0["reg_x"], d0["reg_u"], np.zeros(B * M))
    ds.set_dubins(p)

    def device_step():
        ds.set_warm_start(None, None)  # X_prev = U_prev = 0 (the fixture's start), ADMM/active-set state reset
        ds.scp_solve_async(MAX_IT, RES_TOL)

    # ---- device-resident throughput --------------------------------------------------------------
    for _ in range(W):
        device_step()
    ds.sync()
    barrier()
    launches0 = lib.pmpc_launch_count()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        ds.timer_start()
        kernel_ms = []
        for _ in range(K):
            device_step()
        ms = ds.timer_stop()
        wall = time.perf_counter() - t0
    launches = lib.pmpc_launch_count() - launches0
    barrier()
    ms = parallel.reduce_max(ms, dist, dev_t)
    launches = int(parallel.reduce_sum(launches, dist, dev_t))
    info = ds.info()
    X, U = ds.solution()
    X, U = X.copy(), U.copy()
    hist = ds.hist(int(info["scp_iters"].max()))
    last_res = np.array([hist["resid"][info["scp_iters"][b] - 1, b] for b in range(B)])
    # (a problem that is still above res_tol after max_it iterations stops there, as in the reference: scp_mpc.py:424-428)
    at_max = info["scp_iters"] >= MAX_IT
    ok = bool((info["status"] == 0).all() and ((last_res < RES_TOL) | at_max).all() and np.isfinite(X).all()
              and np.abs(U).max() <= 1 + 1e-9 and np.all(U[:, :, 0] == U[:, :1, 0]))
    kernel_ms = ds.last_kernel_ms()  # the last step's SCP kernel alone
    occ = ds.occupancy()
    value = world * B * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the SCP kernel) -----------------------------------------
    peak, peak_src = hbm_peak()
    abytes, aflops = alg_bytes_flops(info)
    roofline = {"bound": "hbm", "achieved": abytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": abytes / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": None,
                "kernel": ("scp_kernel_half<4,2> (two problems per warp) + resume launch" if args.workload == "c3" else "scp_kernel<4,2,grouped>"),
                "peak_source": peak_src, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": abytes,
                "model": "fused two-sweep pass: 486 B per (particle, stage, pass) = X_prev,U_prev,act in; K,k,H01,sparse stage "
                         "out+in; X,U,act out (DESIGN.md 5)",
                "fp64_tflops_dense_equivalent": aflops / (kernel_ms * 1e-3) / 1e12,
                "unfused_streaming_model_bytes_per_launch": survey_streaming_bytes(info)}
    # the same kernel on SURVEY.md 8(d)'s own denominator (problem data once per SCP iteration, 544 B per unit): what an
    # implementation that kept the whole factorisation on chip would have to move
    sbytes = survey_resident_bytes(info)
    roofline.update(frac_survey=sbytes / (kernel_ms * 1e-3) / 1e9 / peak, survey_bytes_per_launch=sbytes,
                    survey_model="SURVEY 8(d) on-chip-resident case: 8(2n^2+nm+m^2+4n+4m) = 544 B per (particle, stage) per SCP iteration")
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            tr = json.load(fh)
        tr = tr.get(args.workload, tr if args.workload == "c2" else None)
    else:
        tr = None
    if tr:
        # ncu (--set full) DRAM read + write bytes of the same kernel, per problem, scaled to this launch
        roofline["traffic"] = tr["dram_bytes_per_problem"] * B
        roofline["traffic_source"] = tr.get("source")
    if rank == 0:
        for kind, key in ((0, "hbm_copy_gbs_here"), (1, "fp64_fma_tflops_here")):
            r = C.c_double()
            if lib.pmpc_microbench(local, kind, C.byref(r)) == 0:
                roofline[key] = r.value

    # ---- end to end through the public API with host buffers ---------------------------------------
    ua = {k: pmpc_b200.pinned_copy(v) for k, v in untiled(arrs).items()}
    if args.e2e_chunks <= 0:
        args.e2e_chunks = int(max(1, min(16, B // max(1, occ[0] * occ[1]))))

    def e2e_step(solver):
        return pmpc_b200.solve_batch(p, ua["Q"], ua["R"], ua["x0"], ua["X_ref"], ua["U_ref"], ua["X_prev"],
                                     ua["U_prev"], None, None, ua["u_l"], ua["u_u"], max_it=MAX_IT, res_tol=RES_TOL,
                                     reg_x=d0["reg_x"], reg_u=d0["reg_u"], slew_rate=0.0, Nc=NC, device=local, solver=solver,
                                     reuse_buffers=True, chunks=args.e2e_chunks)

    pipe = ds if args.e2e_chunks <= 1 else None
    for _ in range(min(W, 2)):
        Xe, Ue, infoe, pipe = e2e_step(pipe)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        Xe, Ue, infoe, pipe = e2e_step(pipe)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = parallel.reduce_max(e2e_s, dist, dev_t)
    h2d = sum(a.nbytes for a in ua.values()) + p.nbytes + 8 * B * M  # inputs as uploaded + params + slew vector
    d2h = Xe.nbytes + Ue.nbytes + 8 * 8 * B
    e2e = {"value": world * B * K / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * e2e_s / K, "api": "pmpc_b200.solve_batch (page-locked host numpy arrays in, page-locked host arrays out; the batch is "
                  "pipelined in %d sub-batches so that uploads / downloads overlap the solve kernels; inputs that are "
                  "identical for all particles are passed once and expanded on the device)" % args.e2e_chunks}
    ok = ok and np.array_equal(Xe[:, :, 1:], X.reshape(Xe[:, :, 1:].shape)) and bool((infoe["status"] == 0).all())

    ref_table = reference_table_check() if rank == 0 else None
    if ref_table is not None and ref_table["ok"] is False:
        ok = False

    # ---- single-solve latency through pmpc_b200.solve (the call a reference user makes) -----------
    lat = []
    if rank == 0 and args.latency_solves > 0:
        car = pmpc_b200.DubinsCar(p[0])
        kw = dict(X_ref=arrs["X_ref"][0], U_ref=arrs["U_ref"][0], X_prev=arrs["X_prev"][0], U_prev=arrs["U_prev"][0],
                  u_l=arrs["u_l"][0], u_u=arrs["u_u"][0], reg_x=d0["reg_x"], reg_u=d0["reg_u"], max_it=MAX_IT,
                  res_tol=RES_TOL, solver_settings=dict(d0["solver_settings"]))
        for i in range(args.latency_solves + 3):
            t0 = time.perf_counter()
            X1, U1, data1 = pmpc_b200.solve(car, arrs["Q"][0], arrs["R"][0], arrs["x0"][0], **kw)
            dt = time.perf_counter() - t0
            if i >= 3:
                lat.append(dt)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fx = c2_iters_fixture()
        its, cores = args.cpu_sample_its, os.cpu_count() or 1
        if args.workload == "c3":
            v, sample = cpu_whole_solves(cores, 3)
            cpu = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample}
        else:
            sec = cpu_sample(cores, its)
        if cpu is None:
            cpu = {"value": cores / (sec * fx["c_port_scp_iters"] / its), "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"{cores} host threads, each the first {its} of {fx['c_port_scp_iters']} SCP iterations of its own C2 "
                         f"problem with the reference's solver restated in C (oracle/c/pmpc_oracle.c, OSQP 0.6.2 defaults), "
                         f"{sec:.1f} s wall, scaled to full solves"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": ("strong" if args.workload == "c3" and args.batch <= 0 else "weak"),
                "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": config_dict(),
                "details": {"problems_per_gpu_per_step": B, "ctas_per_sm": occ[0], "sms": occ[1],
                            "working_set_gb_per_gpu": B * M * N * 135 * 8 / 1e9, "parallelism": "problem batch sharded, no collective",
                            "julia_on_this_box": __import__("shutil").which("julia")},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "parity_ok": ok, "reference_table": ref_table,
                "scp_iters": {"min": int(info["scp_iters"].min()), "max": int(info["scp_iters"].max()), "mean": float(info["scp_iters"].mean()),
                              "stopped_at_max_it": int(at_max.sum()),
                              "problem0": int(info["scp_iters"][0]), "reference_port": (c2_iters_fixture() if args.workload == "c2" else None)},
                "qp_passes_per_solve": {"factorizations": float(info["factorizations"].mean()),
                                        "polish": float(info["polish_passes"].mean()), "admm": float(info["admm_iters"].mean())},
                "latency_ms": ({"p50": 1e3 * float(np.median(lat)), "min": 1e3 * float(min(lat)), "n": len(lat),
                                "api": "plain pmpc_b200.solve(...) call, one problem, host arrays (handles are cached inside solve)"} if lat else None),
                "wall_ms_per_step": 1e3 * wall / K}
        _emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0 if ok else 1


def _emit(line: dict):
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner to stdout) are kept off it
    by pointing fd 1 at stderr for the whole run (see __main__) and writing the result to the saved descriptor."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


_STDOUT_FD = 1
if __name__ == "__main__":
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main())

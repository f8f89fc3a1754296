"""Throughput of the task farm over the wire (pmpc_b200.remote): K problems of BASELINE config 3's shape (N=30, M=16,
Nc=1, Dubins particles) sent by ONE client process through the reference's scheduler (`remote.solve_problems`, one job per
worker address in flight) to ONE GPU worker process that answers on `slots` ports and coalesces the waiting requests into
batched launches.  Prints one JSON line per slot count, with the in-process batch (`pmpc_b200.solve_problems`) beside
it.  The wire numbers are bounded by the client's and worker's Python (pickle + zstd + ZeroMQ per problem), not by
the GPU: they say what an unchanged farm client gets, the device-resident figures are bench.py's.

    python scripts/remote_farm_bench.py [K] [slots ...]        (FARM_WINDOW_MS=<ms>[,<ms>...]: the worker's collecting window)
"""
import json
import os
import socket
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import pmpc_b200  # noqa: E402
from pmpc_b200 import remote  # noqa: E402


def problems(K, N=30, M=16, seed=2):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(K):
        car = pmpc_b200.DubinsCar(p=np.column_stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)]))
        out.append(dict(f_fx_fu_fn=car, Q=np.tile(np.eye(4), (M, N, 1, 1)), R=np.tile(1e-2 * np.eye(2), (M, N, 1, 1)),
                        x0=np.tile(1.0 + 0.1 * rng.standard_normal(4), (M, 1)),
                        X_ref=np.tile(0.1 * rng.standard_normal((N, 4)), (M, 1, 1)), u_l=-np.ones((M, N, 2)),
                        u_u=np.ones((M, N, 2)), reg_x=1.0, reg_u=1e-2, max_it=100, res_tol=1e-5, verbose=False,
                        solver_settings=dict(Nc=1, solver="osqp")))
    return out


def free_port_run(n):
    """n consecutive ports, all free right now on the loopback interface."""
    for _ in range(500):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            base = s.getsockname()[1]
        if base + n >= 65535:
            continue
        held = []
        try:
            for i in range(n):
                held.append(socket.socket())
                held[-1].bind(("127.0.0.1", base + i))
            return base
        except OSError:
            continue
        finally:
            for h in held:
                h.close()
    raise RuntimeError("no run of %d free ports" % n)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    slot_counts = [int(a) for a in sys.argv[2:]] or [16, 128]
    probs = problems(K)
    pmpc_b200.solve_problems(probs[:64])  # context, module load
    t = time.time()
    local = pmpc_b200.solve_problems(probs)
    t_local = time.time() - t
    print(json.dumps(dict(path="in-process batch", problems=K, seconds=round(t_local, 4), solves_per_s=round(K / t_local, 1))), flush=True)
    windows = [float(w) for w in os.getenv("FARM_WINDOW_MS", "1").split(",")]
    for slots in slot_counts:
        for window_ms in windows:
            port = free_port_run(slots)
            srv = remote.Server(port, None, slots=slots, batch_window=1e-3 * window_ms)
            try:
                assert srv.wait_ready(300.0), "worker did not come up"
                remote.WORKER_ADDRS[:] = [("127.0.0.1", port + i) for i in range(slots)]
                remote.solve_problems(probs[: 2 * slots], rescan=False, max_solve_time=120.0)  # warm
                before = remote.call("server_info", "127.0.0.1", port, True)
                t = time.time()
                out = remote.solve_problems(probs, rescan=False, max_solve_time=120.0)
                dt = time.time() - t
                info = remote.call("server_info", "127.0.0.1", port, True)
                worst = max(float(np.abs(a[1] - b[1]).max()) for a, b in zip(out, local))
                print(json.dumps(dict(path="wire, 1 client process -> 1 GPU worker", slots=slots, window_ms=window_ms, problems=K,
                                      seconds=round(dt, 4), solves_per_s=round(K / dt, 1),
                                      worker_cycles=info["cycles"] - before["cycles"],
                                      coalesced_calls=info["coalesced_calls"] - before["coalesced_calls"],
                                      largest_batch=info["largest_batch"], max_abs_diff_vs_in_process=worst)), flush=True)
            finally:
                srv.stop()


if __name__ == "__main__":
    main()

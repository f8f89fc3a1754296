"""pmpc_b200.remote: the reference's task-farm wire protocol (pmpc/remote.py), served by a GPU worker.

CPU tests cover the bytes (zstd frames, request / reply layout), the serving loop with stand-in methods (answers, error
strings, coalescing of waiting `solve` requests into one batch call, isolation of a bad request), the client scheduler
(order, None entries, a worker that never answers), a real spawned worker process on a box without a device (it must
answer a solve request with the library's loud failure, not with numbers), and the UNMODIFIED reference client from
baseline/_ref talking to this worker.  The GPU test runs real solves through the wire and compares them with the in-process
calls (single calls bit for bit; farmed problems, which the worker batches differently, to 1e-12)."""
import inspect
import os
import pickle
import socket
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

from pmpc_b200 import remote

zmq = pytest.importorskip("zmq")
cloudpickle = pytest.importorskip("cloudpickle")

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "baseline", "_ref")
HOST = "127.0.0.1"


def free_ports(n):
    """n consecutive TCP ports that are free right now on the loopback interface."""
    for _ in range(200):
        with socket.socket() as s:
            s.bind((HOST, 0))
            base = s.getsockname()[1]
        if base + n >= 65535:
            continue
        socks = []
        try:
            for i in range(n):
                t = socket.socket()
                socks.append(t)
                t.bind((HOST, base + i))
            return base
        except OSError:
            continue
        finally:
            for t in socks:
                t.close()
    pytest.skip("no run of %d free ports" % n)


class Hosted:
    """A Worker served from a thread of the test process (stand-in methods; no device)."""

    def __init__(self, slots=1, methods=None, batch_solve=None, batch_window=0.05):
        self.port = free_ports(slots)
        self.worker = remote.Worker(self.port, slots=slots, methods=methods, batch_solve=batch_solve,
                                    batch_window=batch_window).bind()
        self.stop = threading.Event()
        self.thread = None

    def start(self):
        def loop():
            while not self.stop.is_set():
                self.worker.serve_once(20)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()
        return self

    def close(self):
        self.stop.set()
        if self.thread is not None:
            self.thread.join(5)
        self.worker.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def fake_solve(f_fx_fu_fn, Q, R, x0, *args, **kw):
    """Deterministic stand-in with the result layout of solve: X (N+1, n), U (N, m), data."""
    if kw.get("explode", False):
        raise ValueError("bad problem %r" % (kw["explode"],))
    N, n, m = np.shape(Q)[-3], np.shape(Q)[-1], np.shape(R)[-1]
    X = np.concatenate([np.reshape(x0, (1, n)), np.full((N, n), float(np.sum(x0)))])
    return X, np.full((N, m), float(kw.get("reg_x", 1.0))), dict(hist=[dict(it=1, resid=0.0)], tag=f_fx_fu_fn)


def problem(i, N=5, n=3, m=2, **kw):
    return dict(f_fx_fu_fn="dyn%d" % i, Q=np.tile(np.eye(n), (N, 1, 1)), R=np.tile(np.eye(m), (N, 1, 1)),
                x0=np.arange(n) + float(i), reg_x=1.0 + i, verbose=False, **kw)


def wait_all(polls, timeout=20.0):
    out, t0 = [remote.NOT_ARRIVED_YET] * len(polls), time.time()
    while any(isinstance(o, str) and o == remote.NOT_ARRIVED_YET for o in out):
        assert time.time() - t0 < timeout, "no reply within %.0f s" % timeout
        for i, p in enumerate(polls):
            if isinstance(out[i], str) and out[i] == remote.NOT_ARRIVED_YET:
                out[i] = p()
        time.sleep(1e-3)
    return out


# -- bytes -----------------------------------------------------------------------------------------
def test_libzstd_shim_writes_and_reads_standard_frames():
    """The ctypes codec used when the `zstandard` package is absent: round trip, the frame magic, interchange with an
    independent zstd implementation (pyarrow's bundled codec), and ZstdError on garbage / truncation."""
    z = remote._LibZstd()
    rng = np.random.default_rng(0)
    for blob in (b"", b"x", bytes(rng.integers(0, 4, 100_000, dtype=np.uint8)), rng.standard_normal(5000).tobytes()):
        c = z.compress(blob)
        assert c[:4] == b"\x28\xb5\x2f\xfd" and z.decompress(c) == blob
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("zstd")
    blob = bytes(rng.integers(0, 7, 50_000, dtype=np.uint8))
    assert z.decompress(codec.compress(blob, asbytes=True)) == blob
    assert codec.decompress(z.compress(blob), decompressed_size=len(blob), asbytes=True) == blob
    with pytest.raises(z.ZstdError):
        z.decompress(b"definitely not a frame")
    with pytest.raises(z.ZstdError):
        z.decompress(z.compress(blob)[:-7])


def test_request_and_reply_layout_follow_the_reference():
    """remote.py:72-74 / 240-250 / 268-269: request = pickle((sys.path, zstd(pickle((method, args, kwargs))))), reply =
    zstd(pickle(result)); constants and call attributes a client script reads."""
    msg = remote.pack_request("solve", (1, "a"), dict(k=np.arange(3)))
    syspath, data = pickle.loads(msg)
    assert syspath == sys.path and isinstance(data, bytes)
    method, args, kwargs = pickle.loads(remote.COMPRESSION_MODULE.decompress(data))
    assert method == "solve" and tuple(args) == (1, "a") and np.array_equal(kwargs["k"], np.arange(3))
    m, a, k = remote.unpack_request(msg)
    assert (m, a) == ("solve", (1, "a")) and np.array_equal(k["k"], np.arange(3))
    ret = (np.ones((2, 2)), None, dict(hist=[1]))
    back = remote.unpack_reply(remote.pack_reply(ret))
    assert np.array_equal(back[0], ret[0]) and back[1:] == ret[1:]
    assert pickle.loads(remote.COMPRESSION_MODULE.decompress(remote.pack_reply("err"))) == "err"
    assert remote.DEFAULT_PORT == 65535 - 7117 and remote.DEFAULT_HOSTNAME == "localhost"
    for f in (remote.solve, remote.tune_scp):
        assert (f.hostname, f.port, f.blocking) == ("localhost", remote.DEFAULT_PORT, True)
    assert list(inspect.signature(remote.call).parameters) == ["method", "hostname", "port", "blocking", "args", "kwargs"]
    assert list(inspect.signature(remote.solve_problems).parameters) == [
        "problems", "verbose", "randomize_assignment", "rescan", "max_solve_time", "redis_config"]
    assert list(inspect.signature(remote.solve_problem).parameters) == ["problem", "solve_fn"]
    assert list(inspect.signature(remote.start_server).parameters)[:3] == ["port", "verbose", "redis_config"]
    for name in ("Server", "RegisteredFunction", "precompilation_call", "rescan_workers", "solve_problem_remote",
                 "get_redis_connection", "set_redis_status", "unset_redis_status", "main", "SERVERS", "WORKER_ADDRS"):
        assert hasattr(remote, name), name


def test_solve_problem_argument_mapping():
    seen = {}
    remote.solve_problem(dict(f_fx_fu_fn="f", Q="Q", R="R", x0="x0", max_it=3), solve_fn=lambda *a, **k: seen.update(a=a, k=k))
    assert seen["a"] == ("f", "Q", "R", "x0") and seen["k"] == dict(max_it=3, verbose=True)
    p = remote._as_problem(("f", "Q", "R", "x0", "Xr"), dict(max_it=7, profile=True))
    assert p == dict(f_fx_fu_fn="f", Q="Q", R="R", x0="x0", X_ref="Xr", max_it=7, profile=True)


def test_builtin_dynamics_travel_without_their_device_handle():
    """Problems reach a worker as pickles: a model that has cached a device handle (after being called) must still
    pickle, as its description only."""
    from pmpc_b200 import DubinsCar, LTI, UserDynamics

    car = DubinsCar(p=np.array([[0.9, 0.8, 0.3], [1.0, 1.0, 0.3]]), eps=1e-3)
    car._solver = threading.Lock()  # something unpicklable, as a ctypes-backed handle is
    back = pickle.loads(cloudpickle.dumps(car))
    assert back._solver is None and back.eps == 1e-3 and np.array_equal(back.p, car.p) and car._solver is not None
    usr = UserDynamics("__device__ void f(){}", "f", 3, 1, params=np.arange(4.0))
    usr._solver = threading.Lock()
    back = pickle.loads(cloudpickle.dumps(usr))
    assert (back.src, back.fn_name, back.xdim, back.udim, back._solver) == (usr.src, "f", 3, 1, None)
    lti = pickle.loads(cloudpickle.dumps(LTI(np.eye(3), np.ones((3, 1)))))
    assert lti.A.shape == (3, 3) and lti.c is None


# -- serving loop ----------------------------------------------------------------------------------
def test_worker_answers_calls_errors_and_garbage():
    """Every request gets a reply: the result, the traceback string of a failing method (remote.py:270-277), a message
    for an unknown method, and an error string for bytes that are not a request at all."""
    with Hosted(methods=dict(solve=fake_solve, echo=lambda *a, **k: (a, k))).start() as h:
        assert remote.call("echo", HOST, h.port, True, 1, 2, z=3) == ((1, 2), dict(z=3))
        p = problem(2)
        X, U, data = remote.solve_problem(p, solve_fn=lambda *a, **k: remote.call("solve", HOST, h.port, True, *a, **k))
        Xe, Ue, de = fake_solve(p["f_fx_fu_fn"], p["Q"], p["R"], p["x0"], reg_x=p["reg_x"])
        assert np.array_equal(X, Xe) and np.array_equal(U, Ue) and data == de
        err = remote.call("solve", HOST, h.port, True, "f", p["Q"], p["R"], p["x0"], explode=17)
        assert isinstance(err, str) and "ValueError: bad problem 17" in err and "Traceback" in err
        err = remote.call("rm_rf", HOST, h.port, True)
        assert isinstance(err, str) and "Unsupported method 'rm_rf'" in err
        info = remote.call("server_info", HOST, h.port, True)
        assert info["ports"] == [h.port] and info["requests"] >= 4 and info["backend"] == "pmpc_b200"
        # raw garbage, and a well-formed outer pickle around bytes that are not zstd
        for junk in (b"\x00\x01 not a pickle", cloudpickle.dumps((sys.path, b"not zstd"))):
            s = zmq.Context.instance().socket(zmq.REQ)
            s.setsockopt(zmq.LINGER, 0)
            s.connect("tcp://%s:%d" % (HOST, h.port))
            s.send(junk)
            assert s.poll(5000), "the worker must answer undecodable requests too"
            assert "Traceback" in remote.unpack_reply(s.recv())
            s.close(0)
        assert remote.call("echo", HOST, h.port, True, "still alive") == (("still alive",), {})
        # remote.solve / remote.tune_scp go where their attributes point (remote.py:100-120)
        saved = [(f, f.hostname, f.port, f.blocking) for f in (remote.solve, remote.tune_scp)]
        try:
            remote.solve.hostname, remote.solve.port = HOST, h.port
            X2, U2, _ = remote.solve(p["f_fx_fu_fn"], p["Q"], p["R"], p["x0"], reg_x=p["reg_x"])
            assert np.array_equal(X2, Xe) and np.array_equal(U2, Ue)
            remote.tune_scp.hostname, remote.tune_scp.port, remote.tune_scp.blocking = HOST, h.port, False
            assert "Unsupported method 'tune_scp'" in wait_all([remote.tune_scp(1, 2)])[0]
        finally:
            for f, a, b, c in saved:
                f.hostname, f.port, f.blocking = a, b, c


def test_results_leave_the_worker_without_process_local_handles():
    """`solve` parks its device handle in data["device_solver"]; that key stays on the worker.  A result that cannot be
    pickled at all is answered with the pickling traceback."""
    methods = dict(solve=lambda *a, **k: (np.ones(2), np.zeros(1), dict(hist=[1], device_solver=threading.Lock())),
                   tune_scp=lambda *a, **k: threading.Lock())
    with Hosted(methods=methods).start() as h:
        X, U, data = remote.call("solve", HOST, h.port, True)
        assert np.array_equal(X, np.ones(2)) and data == dict(hist=[1])
        err = remote.call("tune_scp", HOST, h.port, True)
        assert isinstance(err, str) and "Traceback" in err and "pickle" in err.lower()


def test_waiting_solve_requests_become_one_batch_call():
    """K solve requests that are waiting when the worker looks (K ports of one ROUTER socket) reach `batch_solve` as ONE
    list, in the keyword form solve_problems takes; other methods in the same drain are served individually; replies go
    back to the right clients."""
    K, batches = 6, []

    def batch_solve(problems):
        batches.append([p["f_fx_fu_fn"] for p in problems])
        assert all(set(p) >= {"f_fx_fu_fn", "Q", "R", "x0", "reg_x", "verbose"} for p in problems)
        return [fake_solve(**p) for p in problems]

    with Hosted(slots=K, methods=dict(solve=fake_solve, echo=lambda x: x), batch_solve=batch_solve, batch_window=0.2) as h:
        assert h.worker.ports == [h.port + i for i in range(K)]
        probs = [problem(i) for i in range(K)]
        polls = [remote.solve_problem_remote((probs[i], HOST, h.port + i, False)) for i in range(K)]
        polls.append(remote.call("echo", HOST, h.port, False, "hello"))
        time.sleep(0.5)  # all seven requests are in the socket's queue now
        assert all(p() == remote.NOT_ARRIVED_YET for p in polls)
        served = 0
        while served < K + 1:
            served += h.worker.serve_once(1000)
        out = wait_all(polls)
        assert out[-1] == "hello"
        assert len(batches) == 1 and sorted(batches[0]) == ["dyn%d" % i for i in range(K)]
        for i in range(K):
            Xe, Ue, de = fake_solve(**probs[i])
            assert np.array_equal(out[i][0], Xe) and np.array_equal(out[i][1], Ue) and out[i][2]["tag"] == "dyn%d" % i
        assert (h.worker.coalesced_calls, h.worker.largest_batch, h.worker.requests) == (1, K, K + 1)
        # a single waiting request is served by `solve` itself, not through the batch path
        poll = remote.solve_problem_remote((probs[0], HOST, h.port, False))
        time.sleep(0.2)
        assert h.worker.serve_once(1000) == 1 and len(batches) == 1
        assert np.array_equal(wait_all([poll])[0][1], fake_solve(**probs[0])[1])


def test_one_bad_request_does_not_fail_its_batch():
    """If the batch call raises, every request of the drain is answered on its own: the bad one gets its traceback, the
    others their solutions."""
    def batch_solve(problems):
        if any(p.get("explode") for p in problems):
            raise ValueError("batch refused")
        return [fake_solve(**p) for p in problems]

    with Hosted(slots=3, methods=dict(solve=fake_solve), batch_solve=batch_solve, batch_window=0.2) as h:
        probs = [problem(0), problem(1, explode="yes"), problem(2)]
        polls = [remote.solve_problem_remote((probs[i], HOST, h.port + i, False)) for i in range(3)]
        time.sleep(0.4)
        served = 0
        while served < 3:
            served += h.worker.serve_once(1000)
        out = wait_all(polls)
        assert isinstance(out[1], str) and "bad problem 'yes'" in out[1]
        for i in (0, 2):
            assert np.array_equal(out[i][0], fake_solve(**probs[i])[0])
        assert h.worker.coalesced_calls == 0


# -- client scheduler ------------------------------------------------------------------------------
@pytest.fixture
def worker_addrs():
    saved = list(remote.WORKER_ADDRS)
    yield remote.WORKER_ADDRS
    remote.WORKER_ADDRS[:] = saved


def test_solve_problems_over_the_wire_keeps_order_and_none_entries(worker_addrs):
    """remote.py:391-452: results in the order of the problems, None for a None problem, one job per worker address."""
    with Hosted(slots=4, methods=dict(solve=fake_solve), batch_solve=lambda ps: [fake_solve(**p) for p in ps],
                batch_window=2e-3).start() as h:
        worker_addrs[:] = [(HOST, h.port + i) for i in range(4)]
        probs = [None if i % 5 == 3 else problem(i) for i in range(23)]
        out = remote.solve_problems(probs, rescan=False, max_solve_time=10.0)
        assert len(out) == 23
        for i, (p, r) in enumerate(zip(probs, out)):
            if p is None:
                assert r is None
            else:
                Xe, Ue, _ = fake_solve(**p)
                assert np.array_equal(r[0], Xe) and np.array_equal(r[1], Ue) and r[2]["tag"] == "dyn%d" % i
        assert h.worker.requests == sum(p is not None for p in probs)
        out2 = remote.solve_problems(probs[:7], rescan=False, randomize_assignment=False)
        assert all((a is None and b is None) or np.array_equal(a[1], b[1]) for a, b in zip(out2, out[:7]))


def test_a_worker_that_never_answers_is_dropped_and_its_job_requeued(worker_addrs, capsys):
    """remote.py:426-434: a job held longer than max_solve_time goes back to the queue and the worker is not used again;
    with no worker left the call fails loudly instead of spinning forever."""
    mute_port = free_ports(1)
    mute = zmq.Context.instance().socket(zmq.ROUTER)  # accepts requests, never replies
    mute.setsockopt(zmq.LINGER, 0)
    mute.bind("tcp://%s:%d" % (HOST, mute_port))
    try:
        with Hosted(slots=1, methods=dict(solve=fake_solve)).start() as h:
            worker_addrs[:] = [(HOST, mute_port), (HOST, h.port)]
            probs = [problem(i) for i in range(6)]
            out = remote.solve_problems(probs, rescan=False, max_solve_time=0.5)
            assert all(np.array_equal(r[1], fake_solve(**p)[1]) for p, r in zip(probs, out))
            assert "is broken, we will use another worker" in capsys.readouterr().out
            assert h.worker.requests == 6
        worker_addrs[:] = [(HOST, mute_port)]
        with pytest.raises(RuntimeError, match="no PMPC worker answers"):
            remote.solve_problems([problem(0)], rescan=False, max_solve_time=0.3)
    finally:
        mute.close(0)


def test_without_workers_the_local_batch_path_is_taken(worker_addrs, monkeypatch):
    """No worker address anywhere: the reference waits forever; here the list goes to pmpc_b200.solve_problems in this
    process (None entries stay None, `verbose` is not forwarded into the problems)."""
    worker_addrs[:] = []
    assert remote.solve_problems([None, None]) == [None, None]
    seen = {}

    def local(problems, verbose=False):
        seen["p"] = problems
        return ["r%d" % i for i in range(len(problems))]

    monkeypatch.setattr(remote.scp_mpc, "solve_problems", local)
    assert remote.solve_problems([problem(0), None, problem(2)]) == ["r0", None, "r1"]
    assert [p["f_fx_fu_fn"] for p in seen["p"]] == ["dyn0", "dyn2"] and all(p["verbose"] is False for p in seen["p"])


# -- worker process --------------------------------------------------------------------------------
def test_spawned_worker_on_a_box_without_a_device_fails_loudly():
    """A real worker process (spawned, heartbeat, stop).  Without a CUDA device: the warm-up call fails and the process
    reports it (no silent CPU service); with the warm-up skipped, a solve request is answered with the library's error
    traceback — never with numbers."""
    from pmpc_b200 import _lib

    try:
        have_gpu = _lib.device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        pytest.skip("this test is about a box without a device")
    port = free_ports(2)
    srv = remote.Server(port, None, precompile=False, slots=2)
    try:
        assert srv.wait_ready(60.0), "worker process did not come up"
        assert srv.is_alive() and not srv.failed()
        info = remote.call("server_info", HOST, port + 1, True)
        assert info["ports"] == [port, port + 1] and info["pid"] != os.getpid()
        from pmpc_b200 import LTI

        p = dict(f_fx_fu_fn=LTI(0.5 * np.eye(3), np.ones((3, 1))), Q=np.tile(np.eye(3), (4, 1, 1)),
                 R=np.tile(np.eye(1), (4, 1, 1)), x0=np.ones(3), verbose=False, max_it=3)
        ret = remote.solve_problem(p, solve_fn=lambda *a, **k: remote.call("solve", HOST, port, True, *a, **k))
        assert isinstance(ret, str) and "Traceback" in ret and ("PmpcError" in ret or "CUDA" in ret or "libpmpc_b200" in ret)
    finally:
        srv.stop()
    assert srv.process is None
    cold = remote.Server(port, None, slots=1)  # default: warm-up on; must give up, not serve
    try:
        t0 = time.time()
        while not cold.failed() and time.time() - t0 < 60.0:
            time.sleep(0.05)
        assert cold.failed() and not cold.is_alive()
    finally:
        cold.stop()


def test_worker_launcher_refuses_to_start_without_a_device():
    """`python -m pmpc_b200.remote` (the counterpart of `python -m pmpc.remote`, remote.py:456-517) on a box without a
    CUDA device: exits with the library's error instead of starting workers that could not serve."""
    from pmpc_b200 import _lib

    try:
        have_gpu = _lib.device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        pytest.skip("this test is about a box without a device")
    r = subprocess.run([sys.executable, "-m", "pmpc_b200.remote", "--port", str(free_ports(1)), "--worker-num", "0"],
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, cwd="/tmp", timeout=120)
    assert r.returncode != 0 and ("no CUDA device" in r.stderr or "PmpcError" in r.stderr), r.stderr[-2000:]


REF_CLIENT = r"""
import sys, json, numpy as np
import pmpc.remote as ref                       # the UNMODIFIED reference client (PYTHONPATH = baseline/_ref)
assert "baseline/_ref" in ref.__file__.replace("\\", "/")
host, port, slots = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
N, n, m = 5, 3, 2
def problem(i):
    return dict(f_fx_fu_fn="dyn%d" % i, Q=np.tile(np.eye(n), (N, 1, 1)), R=np.tile(np.eye(m), (N, 1, 1)),
                x0=np.arange(n) + float(i), reg_x=1.0 + i, verbose=False)
X, U, data = ref.call("solve", host, port, True, "dynA", problem(4)["Q"], problem(4)["R"], problem(4)["x0"], reg_x=5.0)
out = dict(single=[X.tolist(), U.tolist(), data["tag"]])
out["unsupported"] = ref.call("no_such_method", host, port, True)
ref.WORKER_ADDRS[:] = [(host, port + i) for i in range(slots)]
ref.redis = None                                # no registry: use the addresses given above
res = ref.solve_problems([problem(i) for i in range(11)] + [None], rescan=False)
out["farm"] = [None if r is None else [r[0].tolist(), r[1].tolist(), r[2]["tag"]] for r in res]
print(json.dumps(out))
"""


def test_reference_client_unmodified_talks_to_this_worker(tmp_path):
    """Wire compatibility from the other side: the reference's own `pmpc.remote.call` and `pmpc.remote.solve_problems`
    (baseline/_ref, unmodified; its `import zstandard` served by tests/stubs over libzstd) against this worker — single
    call, unsupported method, and its job scheduler over the worker's slot ports."""
    if not os.path.isdir(os.path.join(REF, "pmpc")):
        pytest.skip("baseline/_ref not built (oracle/ref_install.sh needs /root/reference)")
    import json

    with Hosted(slots=4, methods=dict(solve=fake_solve), batch_solve=lambda ps: [fake_solve(**p) for p in ps],
                batch_window=2e-3).start() as h:
        script = tmp_path / "ref_client.py"
        script.write_text(REF_CLIENT)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, os.path.join(ROOT, "tests", "stubs"), ROOT]))
        r = subprocess.run([sys.executable, str(script), HOST, str(h.port), "4"], env=env, capture_output=True, text=True,
                           cwd=str(tmp_path), timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        out = json.loads(r.stdout.strip().splitlines()[-1])
        Xe, Ue, _ = fake_solve(**dict(problem(4), f_fx_fu_fn="dynA", reg_x=5.0))
        assert np.array_equal(out["single"][0], Xe) and np.array_equal(out["single"][1], Ue) and out["single"][2] == "dynA"
        assert "Unsupported method" in out["unsupported"]
        assert out["farm"][-1] is None
        for i in range(11):
            Xe, Ue, _ = fake_solve(**problem(i))
            assert np.array_equal(out["farm"][i][0], Xe) and np.array_equal(out["farm"][i][1], Ue) and out["farm"][i][2] == "dyn%d" % i
        assert h.worker.requests == 13


# -- GPU -------------------------------------------------------------------------------------------
def solved(ret):
    assert not isinstance(ret, str), "the worker answered with an error:\n" + ret
    return ret


def _dubins_problems(K, rng, N=12, M=4):
    from pmpc_b200 import DubinsCar

    out = []
    for _ in range(K):
        car = DubinsCar(p=np.column_stack([rng.uniform(0.7, 1.0, M), rng.uniform(0.7, 1.0, M), np.full(M, 0.3)]), eps=1e-6)
        out.append(dict(f_fx_fu_fn=car, Q=np.tile(np.eye(4), (M, N, 1, 1)), R=np.tile(1e-2 * np.eye(2), (M, N, 1, 1)),
                        x0=np.tile(1.0 + 0.1 * rng.standard_normal(4), (M, 1)), X_ref=np.tile(0.1 * rng.standard_normal((N, 4)), (M, 1, 1)),
                        u_l=-np.ones((M, N, 2)), u_u=np.ones((M, N, 2)), reg_x=1.0, reg_u=1e-2, max_it=30, res_tol=1e-6,
                        verbose=False, solver_settings=dict(Nc=1, solver="osqp")))
    return out


@pytest.mark.gpu
def test_gpu_worker_serves_solves_bit_identical_to_in_process_calls(gpu, worker_addrs):
    """A real worker process on the GPU (spawned, warmed up): one blocking `remote.solve` with a Python f_fx_fu_fn (host
    loop on the worker), one with device dynamics, then a farm of 24 problems over 8 slot ports through
    `remote.solve_problems` — single calls equal bit for bit to the in-process call, farmed ones to 1e-12; the worker reports that it
    coalesced waiting requests into batched launches."""
    import pmpc_b200
    from fixtures import config_c1
    from oracle import dubins

    slots = 8
    port = free_ports(slots)
    srv = remote.Server(port, None, slots=slots)
    try:
        assert srv.wait_ready(300.0), "GPU worker did not come up (warm-up failed?)"
        d, p = config_c1()
        kw = dict(d)
        Q, R, x0 = kw.pop("Q"), kw.pop("R"), kw.pop("x0")
        remote.solve.hostname, remote.solve.port, remote.solve.blocking = HOST, port, False  # polled: a dead worker fails the test, not hangs it
        try:
            fn = dubins.make_f_fx_fu_fn(p)
            Xw, Uw, dw = solved(wait_all([remote.solve(fn, Q, R, x0, **kw)], 180.0)[0])
            Xl, Ul, dl = pmpc_b200.solve(fn, Q, R, x0, **kw)
            assert np.array_equal(Xw, Xl) and np.array_equal(Uw, Ul)
            assert [h["obj"] for h in dw["hist"]] == [h["obj"] for h in dl["hist"]]
            car = pmpc_b200.DubinsCar(p=p, eps=1e-6)
            Xw, Uw, dw = solved(wait_all([remote.solve(car, Q, R, x0, **kw)], 180.0)[0])
            assert "device_solver" not in dw and dw["hist"]
            Xl, Ul, dl = pmpc_b200.solve(car, Q, R, x0, **kw)
            assert np.array_equal(Xw, Xl) and np.array_equal(Uw, Ul)
        finally:
            remote.solve.hostname, remote.solve.port, remote.solve.blocking = remote.DEFAULT_HOSTNAME, remote.DEFAULT_PORT, True
        probs = _dubins_problems(24, np.random.default_rng(5)) + [None]
        worker_addrs[:] = [(HOST, port + i) for i in range(slots)]
        out = remote.solve_problems(probs, rescan=False, max_solve_time=120.0)
        local = pmpc_b200.solve_problems(probs[:-1])
        assert out[-1] is None
        exact = 0
        for (Xw, Uw, dw), (Xl, Ul, dl) in zip(map(solved, out[:-1]), local):
            # the worker's launches group the problems differently from the one local batch: same iterates (the
            # kernels are batch-composition independent, test_batched_equals_individual), compared to 1e-12 here
            assert np.abs(Xw - Xl).max() <= 1e-12 and np.abs(Uw - Ul).max() <= 1e-12
            assert len(dw["hist"]) == len(dl["hist"]) and abs(dw["hist"][-1]["obj"] - dl["hist"][-1]["obj"]) <= 1e-9
            exact += int(np.array_equal(Xw, Xl) and np.array_equal(Uw, Ul))
        print("wire farm: %d of %d results bit-identical to the local batch" % (exact, len(local)))
        info = wait_all([remote.call("server_info", HOST, port, False)], 30.0)[0]
        assert info["requests"] >= 27 and info["pid"] != os.getpid()
        assert info["coalesced_calls"] >= 1 and info["largest_batch"] >= 2, info
    finally:
        srv.stop()


@pytest.mark.gpu
def test_repeated_batches_reuse_one_handle_without_carrying_state(gpu):
    """solve_problems keeps the handle of a moderate batch in solve's cache (a control loop, or a farm worker, sends the
    same batch shape again and again).  A second batch of the same shape but other data, no control bounds and a slew
    penalty, solved on the re-used handle, equals the same batch solved on a fresh handle bit for bit."""
    import pmpc_b200
    from pmpc_b200 import scp_mpc

    A = _dubins_problems(6, np.random.default_rng(11))
    B = _dubins_problems(6, np.random.default_rng(12))
    for p in B:
        del p["u_l"], p["u_u"]
        p.update(slew_rate=0.3, X_prev=0.05 * np.ones_like(p["X_ref"]))
    scp_mpc.clear_handle_cache()
    pmpc_b200.solve_problems(A)
    keys = [k for k in scp_mpc._HANDLE_CACHE if len(k) == 7 and k[-1] == 6]
    assert len(keys) == 1, list(scp_mpc._HANDLE_CACHE)
    handle = scp_mpc._HANDLE_CACHE[keys[0]]
    reused = pmpc_b200.solve_problems(B)
    assert scp_mpc._HANDLE_CACHE[keys[0]] is handle
    scp_mpc.clear_handle_cache()
    fresh = pmpc_b200.solve_problems(B)
    for (Xr, Ur, dr), (Xf, Uf, df) in zip(reused, fresh):
        assert np.array_equal(Xr, Xf) and np.array_equal(Ur, Uf) and len(dr["hist"]) == len(df["hist"])
    single = pmpc_b200.solve(**B[2])
    assert np.array_equal(single[0], fresh[2][0]) and np.array_equal(single[1], fresh[2][1])


@pytest.mark.gpu
def test_worker_launcher_serves_on_every_slot_and_stops_on_sigint(gpu, tmp_path):
    """`python -m pmpc_b200.remote --worker-num 0 --slots 3` (remote.py:456-517 with GPUs in the place of cores): one worker
    per visible GPU comes up warmed, owns its GPU, answers on all of its slot ports, solves, and the launcher stops its
    workers and exits cleanly on SIGINT."""
    import signal

    import pmpc_b200

    slots = 3
    workers = "0" if gpu == 1 else "2"  # 0 = one per visible GPU; on a multi-GPU box two workers are enough for the test
    port = free_ports(slots * min(gpu, 2))
    log = tmp_path / "launcher.log"
    out = lambda: log.read_text()[-3000:]
    with open(log, "w") as fh:
        proc = subprocess.Popen([sys.executable, "-m", "pmpc_b200.remote", "--port", str(port), "--worker-num", workers, "--slots",
                                 str(slots)], env=dict(os.environ, PYTHONPATH=ROOT, PYTHONUNBUFFERED="1"), cwd="/tmp", stdout=fh,
                                stderr=subprocess.STDOUT)
    try:
        info, t0 = None, time.time()
        while info is None and time.time() - t0 < 240.0:
            assert proc.poll() is None, "launcher exited: " + out()
            poll = remote.call("server_info", HOST, port + slots - 1, False)
            try:
                info = wait_all([poll], 2.0)[0]
            except AssertionError:
                poll.ctx.destroy()
        assert info is not None, "no worker answered within 240 s: " + out()
        assert info["ports"] == [port + i for i in range(slots)] and info["device"] == "0" and info["pid"] != proc.pid
        car = pmpc_b200.DubinsCar(p=np.array([1.0, 1.0, 0.3]))
        args = (car, np.tile(np.eye(4), (15, 1, 1)), np.tile(1e-2 * np.eye(2), (15, 1, 1)), np.ones(4))
        kw = dict(u_l=-np.ones((15, 2)), u_u=np.ones((15, 2)), max_it=20, res_tol=1e-6, verbose=False,
                  solver_settings=dict(solver="osqp"))
        Xw, Uw, _ = solved(wait_all([remote.call("solve", HOST, port + 1, False, *args, **kw)], 120.0)[0])
        Xl, Ul, _ = pmpc_b200.solve(*args, **kw)
        assert Xw.shape == (16, 4) and np.array_equal(Xw, Xl) and np.array_equal(Uw, Ul)
        proc.send_signal(signal.SIGINT)
        assert proc.wait(90) == 0, out()
        assert "Stopping server on port %d" % port in out()
    finally:
        if proc.poll() is None:
            proc.kill()
            proc.wait(30)

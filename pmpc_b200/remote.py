"""GPU worker and client for the reference's task-farm wire protocol (``pmpc/remote.py``).

The reference fans independent problems out to worker PROCESSES, one per CPU core, because a Julia/OSQP solve occupies
one core: a ZeroMQ REP socket per worker, requests ``cloudpickle((sys.path, zstd(cloudpickle((method, args, kwargs)))))``,
replies ``zstd(cloudpickle(result))`` (``remote.py:59-97, 209-283``), workers found through redis keys
``pmpc_worker_<host>_<pid>/<host>:<port>`` (``remote.py:186-203, 370-388``).  This module speaks the same bytes, so an
existing client (the reference's own ``pmpc.remote.call`` / ``solve_problems``, unmodified) can point at a B200, and
the functions a client imports from here keep the reference's names, arguments and behaviour.

What is different is the worker, because the unit of work on a B200 is a batch, not a problem:

* ONE worker process per GPU, not per core.  It owns a ROUTER socket (REQ clients cannot tell it from a REP socket)
  bound to ``slots`` consecutive ports, so a client that keeps one job per worker address in flight
  (``remote.py:400-446``) keeps ``slots`` jobs in flight on this GPU.
* Requests that are waiting when the worker comes back from a launch are drained together; the ``solve`` calls among
  them go to ``pmpc_b200.solve_problems``, which turns same-shaped problems with device dynamics into ONE launch (one
  CTA per problem) and solves the rest one by one.  Batched and individual results are bit-identical
  (``tests/test_gpu_parity.py::test_batched_equals_individual``), so coalescing is invisible to the client.
* Every request is answered (the reference drops undecodable ones, which wedges its REP socket): a failure comes back
  as the traceback string, as ``remote.py:263-277`` does for solver errors.

SECURITY.  The protocol is pickles in both directions: whoever can reach a worker's port can run code as the worker.
The reference binds ``tcp://*``.  This worker binds the LOOPBACK interface only (``BIND_HOST``) and there is no switch
for anything else: clients on the same host connect directly, clients on other hosts need a transport the operator
sets up and trusts (an SSH tunnel, ``ssh -L 58418:127.0.0.1:58418 gpu-box``).  The redis registry is read by the
client (``rescan_workers``) but workers register their loopback address, i.e. it serves one host.

zstd: the ``zstandard`` package when importable, else the system's libzstd through ctypes (same frames).
There is no CPU solver behind any of this: a worker on a box without a CUDA device fails its warm-up call and exits.
"""
from __future__ import annotations

import collections
import ctypes
import ctypes.util
import gc
import inspect
import os
import pickle
import random
import signal
import sys
import time
import traceback
from argparse import ArgumentParser
from copy import copy
from socket import gethostbyname, gethostname
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from . import scp_mpc
from .lqp_matrices import lqp_generate_problem_matrices
from .problem_struct import Problem

try:  # both are needed for the wire only
    import zmq
except ModuleNotFoundError:  # pragma: no cover
    zmq = None
try:
    import cloudpickle as serializer
except ModuleNotFoundError:  # pragma: no cover
    serializer = None
try:
    import redis
except ModuleNotFoundError:
    redis = None

DEFAULT_PORT = 65535 - 7117  # remote.py:27
DEFAULT_HOSTNAME = "localhost"
BIND_HOST = "127.0.0.1"  # see SECURITY above
HOSTNAME = gethostname()
PID = os.getpid()
NOT_ARRIVED_YET = "NOT_ARRIVED_YET"  # remote.py:92, what a non-blocking call's poll function returns while waiting
REDIS_CONFIG: Dict[str, Any] = dict()
if os.getenv("REDIS_HOST", None) is not None:
    REDIS_CONFIG["host"] = os.getenv("REDIS_HOST")
if os.getenv("REDIS_PORT", None) is not None:
    REDIS_CONFIG["port"] = int(os.getenv("REDIS_PORT"))
if os.getenv("REDIS_PASSWORD", None) is not None:
    REDIS_CONFIG["password"] = os.getenv("REDIS_PASSWORD")

_POSITIONAL = ("f_fx_fu_fn", "Q", "R", "x0")


# -- zstd ------------------------------------------------------------------------------------------
class _LibZstd:
    """``compress`` / ``decompress`` / ``ZstdError`` of the ``zstandard`` package on top of the system's libzstd
    (one-shot frames with the content size in the header, which is what ``zstandard.compress`` writes and what
    ``zstandard.decompress`` requires)."""

    class ZstdError(Exception):
        pass

    def __init__(self):
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        L = self._L = ctypes.CDLL(name)
        sz, buf, dst = ctypes.c_size_t, ctypes.c_char_p, ctypes.c_void_p
        L.ZSTD_compressBound.restype, L.ZSTD_compressBound.argtypes = sz, [sz]
        L.ZSTD_compress.restype, L.ZSTD_compress.argtypes = sz, [dst, sz, buf, sz, ctypes.c_int]
        L.ZSTD_decompress.restype, L.ZSTD_decompress.argtypes = sz, [dst, sz, buf, sz]
        L.ZSTD_getFrameContentSize.restype, L.ZSTD_getFrameContentSize.argtypes = ctypes.c_ulonglong, [buf, sz]
        L.ZSTD_isError.restype, L.ZSTD_isError.argtypes = ctypes.c_uint, [sz]
        L.ZSTD_getErrorName.restype, L.ZSTD_getErrorName.argtypes = ctypes.c_char_p, [sz]

    def _check(self, rc: int) -> int:
        if self._L.ZSTD_isError(rc):
            raise self.ZstdError(self._L.ZSTD_getErrorName(rc).decode())
        return rc

    def compress(self, data: bytes, level: int = 3) -> bytes:
        data = bytes(data)
        out = ctypes.create_string_buffer(self._L.ZSTD_compressBound(len(data)))
        n = self._check(self._L.ZSTD_compress(out, len(out), data, len(data), level))
        return out.raw[:n]

    def decompress(self, data: bytes) -> bytes:
        data = bytes(data)
        n = self._L.ZSTD_getFrameContentSize(data, len(data))
        if n >= 2 ** 64 - 2:  # ZSTD_CONTENTSIZE_UNKNOWN / ZSTD_CONTENTSIZE_ERROR
            raise self.ZstdError("not a zstd frame with a content size")
        out = ctypes.create_string_buffer(max(int(n), 1))
        if self._check(self._L.ZSTD_decompress(out, len(out), data, len(data))) != n:
            raise self.ZstdError("truncated zstd frame")
        return out.raw[:n]


def _load_compression():
    try:
        import zstandard

        return zstandard
    except ModuleNotFoundError:
        return _LibZstd()


COMPRESSION_MODULE = _load_compression()


# -- wire format -----------------------------------------------------------------------------------
def _need_wire():
    if zmq is None or serializer is None:
        raise ImportError("pmpc_b200.remote: the wire protocol needs pyzmq and cloudpickle")


def pack_request(method: str, args: tuple, kwargs: dict) -> bytes:
    """remote.py:72-74: the client's ``sys.path`` travels along so the worker can import what the pickles refer to."""
    return serializer.dumps((sys.path, COMPRESSION_MODULE.compress(serializer.dumps((method, args, kwargs)))))


def unpack_request(msg: bytes) -> Tuple[str, tuple, dict]:
    """Inverse of :func:`pack_request` (``remote.py:240-250``); extends ``sys.path`` as the reference's worker does."""
    syspath, data = serializer.loads(msg)
    for path in syspath:
        if path not in sys.path:
            sys.path.append(path)
    method, args, kwargs = serializer.loads(COMPRESSION_MODULE.decompress(data))
    return method, tuple(args), dict(kwargs)


def pack_reply(ret: Any) -> bytes:
    """``zstd(pickle(result))``.  Results are plain data (arrays, dicts of numbers, strings): the C pickler writes the
    same stream cloudpickle would, several times faster; cloudpickle is the fallback for anything it refuses."""
    try:
        blob = pickle.dumps(ret, protocol=serializer.DEFAULT_PROTOCOL)
    except Exception:
        blob = serializer.dumps(ret)
    return COMPRESSION_MODULE.compress(blob)


def unpack_reply(msg: bytes) -> Any:
    return serializer.loads(COMPRESSION_MODULE.decompress(msg))


class RegisteredFunction:
    """A function that carries its own pickle as a key, so a worker that has seen it before calls the instance it
    already holds instead of the freshly unpickled one (``remote.py:44-55``)."""

    _seen: Dict[bytes, "RegisteredFunction"] = dict()

    def __init__(self, fn: Callable) -> None:
        self.fn, self.hash = fn, serializer.dumps(fn)

    def __call__(self, *args, **kwargs):
        return self._seen.setdefault(self.hash, self).fn(*args, **kwargs)


# -- client ----------------------------------------------------------------------------------------
class _SocketOwner:
    """Stands where the reference keeps a private ``zmq.Context`` per call (``fn.ctx``): ``destroy`` closes the call's
    socket only, the process-wide context is shared by all calls."""

    def __init__(self, sock):
        self._sock = sock

    def destroy(self, linger=None):
        self._sock.close(0)

    term = destroy


def _request_socket(hostname, port, blocking: bool):
    """A REQ socket to one worker address (``remote.py:76-87``: two-second send / receive limits on non-blocking calls)."""
    _need_wire()
    sock = zmq.Context.instance().socket(zmq.REQ)
    sock.setsockopt(zmq.LINGER, 0)
    if not blocking:
        sock.setsockopt(zmq.RCVTIMEO, 2000)
        sock.setsockopt(zmq.SNDTIMEO, 2000)
    sock.connect("tcp://%s:%s" % (hostname if hostname is not None else DEFAULT_HOSTNAME,
                                  port if port is not None else DEFAULT_PORT))
    return sock


def call(method: str, hostname: Optional[str] = None, port: Optional[int] = None, blocking: bool = True, *args,
         **kwargs) -> Union[Any, Callable]:
    """Call one of the worker's methods (``remote.py:59-97``).  ``blocking``: returns the result.  Otherwise returns a
    poll function that yields ``"NOT_ARRIVED_YET"`` until the reply is there; ``fn.sock`` / ``fn.ctx`` as in the
    reference, for callers that give up on a worker."""
    sock = _request_socket(hostname, port, blocking)
    sock.send(pack_request(method, args, kwargs))
    if blocking:
        try:
            return unpack_reply(sock.recv())
        finally:
            sock.close(0)

    def fn():
        if sock.closed or sock.poll(0) != zmq.POLLIN:
            return NOT_ARRIVED_YET
        ret = unpack_reply(sock.recv())
        sock.close(0)
        return ret

    fn.sock, fn.ctx = sock, _SocketOwner(sock)
    return fn


def solve(*args, **kw):
    """``pmpc.remote.solve`` (``remote.py:100-106``): ``pmpc_b200.solve`` on the worker at ``solve.hostname:solve.port``."""
    return call("solve", solve.hostname, solve.port, solve.blocking, *args, **kw)


def tune_scp(*args, **kw):
    return call("tune_scp", tune_scp.hostname, tune_scp.port, tune_scp.blocking, *args, **kw)


for _f in (solve, tune_scp):
    _f.hostname, _f.port, _f.blocking = DEFAULT_HOSTNAME, DEFAULT_PORT, True


def solve_problem(problem: Dict[str, Any], solve_fn: Callable = solve) -> Tuple[np.ndarray, np.ndarray, Dict[str, Any]]:
    """One problem given as a keyword dictionary (``remote.py:324-343``): ``f_fx_fu_fn, Q, R, x0`` are passed
    positionally, everything else as keywords; ``verbose`` defaults to True as in the reference."""
    rest = {k: v for k, v in problem.items() if k not in _POSITIONAL}
    rest.setdefault("verbose", True)
    if "extra_cstrs_fn_np" in rest:
        rest["extra_cstrs_fn"] = rest["extra_cstrs_fn_np"]
    return solve_fn(*(problem[k] for k in _POSITIONAL), **rest)


def solve_problem_remote(args: Tuple[Dict[str, Any], str, int, bool]):
    """``(problem, hostname, port, blocking)`` -> the result, or the poll function of a non-blocking call
    (``remote.py:346-363``)."""
    problem, hostname, port, blocking = args
    return solve_problem(problem, solve_fn=lambda *a, **k: call("solve", hostname, port, blocking, *a, **k))


def _solve_request(problem: Dict[str, Any]) -> bytes:
    """The bytes ``solve_problem_remote`` sends for this problem."""
    sent = dict()
    solve_problem(problem, solve_fn=lambda *a, **k: sent.update(args=a, kwargs=k))
    return pack_request("solve", sent["args"], sent["kwargs"])


WORKER_ADDRS: List[Tuple[str, Any]] = []
LAST_WORKER_SCAN = time.time() - 60.0


def _redis_connect(redis_config: Optional[Dict[str, Any]]):
    if redis is None:
        return None
    cfg = dict(redis_config if redis_config is not None else REDIS_CONFIG)
    try:
        if "host" in cfg:
            cfg["host"] = gethostbyname(cfg["host"])
        rconn = redis.Redis(**cfg)
        rconn.keys("")  # fails here if the server is not reachable
        return rconn
    except (redis.ConnectionError, OSError):
        print("Could not connect to redis at %s" % (cfg,))
        return None


def rescan_workers(force: bool = False, redis_config: Optional[Dict[str, Any]] = None):
    """Refresh ``WORKER_ADDRS`` from the redis registry, at most every 10 s unless forced (``remote.py:370-388``).
    Without redis the list is whatever the caller put there."""
    global LAST_WORKER_SCAN
    if redis is None or not (force or time.time() - LAST_WORKER_SCAN > 10.0):
        return
    rconn = _redis_connect(redis_config)
    if rconn is None:
        return
    WORKER_ADDRS[:] = [tuple(k.decode("utf8").split("/")[1].split(":")) for k in rconn.keys("pmpc_worker*")]
    if len(WORKER_ADDRS) == 0:
        print("Could not find any PMPC workers registers in redis.")
    LAST_WORKER_SCAN = time.time()


def _solve_in_process(problems, verbose):
    todo = [(i, dict(p, verbose=False)) for i, p in enumerate(problems) if p is not None]
    out: List[Any] = [None] * len(problems)
    for (i, _), res in zip(todo, scp_mpc.solve_problems([p for _, p in todo], verbose=verbose)):
        out[i] = res
    return out


def solve_problems(problems: List[Optional[Dict[str, Any]]], verbose: bool = False, randomize_assignment: bool = True,
                   rescan: bool = True, max_solve_time: float = 20.0, redis_config: Optional[Dict[str, Any]] = None):
    """``pmpc.remote.solve_problems`` (``remote.py:391-452``): solutions in the order of ``problems``, ``None`` for a
    ``None`` entry.  One job at a time per worker address; a worker that holds a job longer than ``max_solve_time`` is
    dropped and its job goes back to the queue; when every worker is gone the registry is scanned again.  Same requests
    on the wire as the reference's scheduler; the sockets are kept per address for the duration of the call (a REQ socket
    costs more to create than a problem costs to pack) and replies are awaited with a poller instead of a sleep loop.

    With no worker address at all (``WORKER_ADDRS`` empty and nothing in redis) the reference waits forever; here the
    problems are solved by THIS process's GPU as one batch (``pmpc_b200.solve_problems``)."""
    if rescan or len(WORKER_ADDRS) == 0:
        rescan_workers(force=False, redis_config=redis_config)
    if len(WORKER_ADDRS) == 0:
        return _solve_in_process(problems, verbose)
    _need_wire()
    results: Dict[int, Any] = {i: None for i, p in enumerate(problems) if p is None}
    queue = collections.deque(i for i, p in enumerate(problems) if p is not None)
    idle: List[Tuple[str, Any]] = list(WORKER_ADDRS)
    socks: Dict[Tuple[str, Any], Any] = dict()  # one REQ socket per worker address, kept for the whole call
    busy: Dict[Any, Tuple[Tuple[str, Any], int, float]] = dict()  # socket -> (address, problem index, send time)
    poller = zmq.Poller()
    bar = None
    if verbose:
        from tqdm import tqdm

        bar = tqdm(total=len(problems), initial=len(results))
    lonely_since = None
    try:
        while len(results) < len(problems):
            if randomize_assignment:
                random.shuffle(idle)
            while idle and queue:
                addr, idx = idle.pop(), queue.pop()
                sock = socks.get(addr)
                if sock is None:
                    sock = socks[addr] = _request_socket(addr[0], addr[1], blocking=False)
                    poller.register(sock, zmq.POLLIN)
                try:
                    sock.send(_solve_request(problems[idx]))
                except zmq.Again:  # the send queue of a worker that is not there is full
                    queue.append(idx)
                    poller.unregister(sock)
                    socks.pop(addr).close(0)
                    continue
                busy[sock] = (addr, idx, time.time())
            for sock, _ in poller.poll(5 if busy else 0):
                if sock in busy:
                    addr, idx, _ = busy.pop(sock)
                    results[idx] = unpack_reply(sock.recv())
                    idle.append(addr)
                    if bar is not None:
                        bar.update(1)
            now = time.time()
            for sock, (addr, idx, t0) in list(busy.items()):
                if now - t0 > max_solve_time:
                    print("PMPC Worker %s is broken, we will use another worker." % (addr,))
                    del busy[sock]
                    poller.unregister(sock)
                    socks.pop(addr).close(0)
                    queue.append(idx)
            if not busy and not idle and queue:  # every worker was dropped
                lonely_since = lonely_since or time.time()
                if time.time() - lonely_since > max(max_solve_time, 1.0):
                    raise RuntimeError("pmpc_b200.remote.solve_problems: no PMPC worker answers (%d problems left)" % len(queue))
                print("All PMPC workers deemed dead, rescanning all PMPC workers.")
                rescan_workers(force=True, redis_config=redis_config)
                idle = list(WORKER_ADDRS) if redis is not None else []
                time.sleep(0.2)
    finally:
        for sock in socks.values():
            sock.close(0)
        if bar is not None:
            bar.close()
    return [results[i] for i in range(len(problems))]


# -- worker ----------------------------------------------------------------------------------------
_SOLVE_SIGNATURE = inspect.signature(scp_mpc.scp_solve)


def _as_problem(args: tuple, kwargs: dict) -> Dict[str, Any]:
    """The keyword form (what ``solve_problems`` takes) of a ``solve(*args, **kwargs)`` request."""
    bound = _SOLVE_SIGNATURE.bind(*args, **kwargs).arguments
    extra = bound.pop("extra_kw", {})
    return dict(bound, **extra)


def _portable(ret):
    """A result leaves the process: the device handle ``solve`` parks in ``data["device_solver"]`` (for callers that
    keep solving the same problem in place) is process-local and stays behind."""
    if isinstance(ret, tuple) and len(ret) == 3 and isinstance(ret[2], dict) and "device_solver" in ret[2]:
        return ret[0], ret[1], {k: v for k, v in ret[2].items() if k != "device_solver"}
    return ret


def _server_info(worker: "Worker") -> Dict[str, Any]:
    return dict(hostname=HOSTNAME, pid=os.getpid(), ports=list(worker.ports), device=os.getenv("CUDA_VISIBLE_DEVICES"),
                requests=worker.requests, cycles=worker.cycles, coalesced_calls=worker.coalesced_calls,
                largest_batch=worker.largest_batch, backend="pmpc_b200")


class Worker:
    """The serving loop of one GPU worker (the reference's ``_server`` body, ``remote.py:226-281``).

    ``methods`` maps the method names a client may call to callables (default: ``solve``, ``tune_scp``,
    ``lqp_generate_problem_matrices`` as ``remote.py:24-26``, plus ``server_info``); ``batch_solve`` takes a list of
    problem dictionaries and returns their results in order (default: ``pmpc_b200.solve_problems``).  ``batch_window``
    [s]: once two requests have been seen together, keep collecting until none arrives for this long."""

    def __init__(self, port: int = DEFAULT_PORT, slots: int = 1, methods: Optional[Dict[str, Callable]] = None,
                 batch_solve: Optional[Callable] = None, batch_window: float = 1e-3, max_batch: int = 1 << 15,
                 verbose: bool = False):
        _need_wire()
        self.ports = [port + i for i in range(max(1, int(slots)))]
        self.methods = dict(methods) if methods is not None else dict(
            solve=scp_mpc.solve, tune_scp=scp_mpc.tune_scp, lqp_generate_problem_matrices=lqp_generate_problem_matrices)
        self.methods.setdefault("server_info", lambda: _server_info(self))
        self.batch_solve = batch_solve if batch_solve is not None or methods is not None else (
            lambda problems: scp_mpc.solve_problems(problems, verbose=False))
        self.batch_window, self.max_batch, self.verbose = float(batch_window), int(max_batch), verbose
        self.requests = self.cycles = self.coalesced_calls = self.largest_batch = 0
        self._crowded = False
        self.sock = None

    def bind(self):
        self.sock = zmq.Context.instance().socket(zmq.ROUTER)
        self.sock.setsockopt(zmq.LINGER, 0)
        try:
            for p in self.ports:
                self.sock.bind("tcp://%s:%d" % (BIND_HOST, p))
        except zmq.ZMQError:
            self.close()
            raise
        return self

    def close(self):
        if self.sock is not None:
            self.sock.close(0)
            self.sock = None

    # one request = (envelope frames up to and including the empty delimiter, payload)
    def _drain(self, timeout_ms: int):
        if not self.sock.poll(timeout_ms):
            return []
        frames, last = [], time.time()
        while len(frames) < self.max_batch:
            try:
                frames.append(self.sock.recv_multipart(zmq.NOBLOCK))
                last = time.time()
            except zmq.Again:
                wait = self.batch_window - (time.time() - last)
                if not (self._crowded or len(frames) > 1) or wait <= 0 or not self.sock.poll(max(1, int(1e3 * wait))):
                    break
        self._crowded = len(frames) > 1
        return frames

    def _run(self, method, args, kwargs):
        if method not in self.methods:
            return "Unsupported method %r; this worker serves %s" % (method, sorted(self.methods))
        try:
            return self.methods[method](*args, **kwargs)
        except Exception:
            err = traceback.format_exc()
            if self.verbose:
                print(err)
            return err

    def serve_once(self, timeout_ms: int = 100) -> int:
        """Answer everything that is waiting (at most ``timeout_ms`` for the first request).  Returns the count."""
        frames = self._drain(timeout_ms)
        if not frames:
            return 0
        replies: List[Any] = [None] * len(frames)
        decoded: Dict[int, Tuple[str, tuple, dict]] = dict()
        for i, fr in enumerate(frames):
            try:
                decoded[i] = unpack_request(fr[-1])
            except Exception:  # not a pickle, not zstd, or a module the pickles need is missing here
                replies[i] = traceback.format_exc()
                if self.verbose:
                    print(replies[i])
        solves = [i for i, (method, _, _) in decoded.items() if method == "solve" and "solve" in self.methods]
        if len(solves) > 1 and self.batch_solve is not None:
            try:
                out = self.batch_solve([_as_problem(*decoded[i][1:]) for i in solves])
                if len(out) != len(solves):
                    raise RuntimeError("batch_solve returned %d results for %d problems" % (len(out), len(solves)))
                for i, r in zip(solves, out):
                    replies[i] = r
                    del decoded[i]
                self.coalesced_calls += 1
                self.largest_batch = max(self.largest_batch, len(solves))
            except Exception:  # one bad request must not fail its neighbours: answer them one by one below
                if self.verbose:
                    print(traceback.format_exc())
        for i, (method, args, kwargs) in decoded.items():
            replies[i] = self._run(method, args, kwargs)
        self.requests += len(frames)  # (before the replies leave: a client may ask for server_info right after its answer)
        self.cycles += 1
        for fr, ret in zip(frames, replies):
            try:
                payload = pack_reply(_portable(ret))
            except Exception:
                payload = pack_reply(traceback.format_exc())
            self.sock.send_multipart(fr[:-1] + [payload])
        return len(frames)


def precompilation_call():
    """Warm-up before the worker starts answering (``remote.py:133-154`` runs small solves to get Julia compiled):
    loads the library, creates the CUDA context and runs the three ways a request can be served — the reference-style
    host loop over a Python ``f_fx_fu_fn`` (``c_lqp_solve`` per iteration), the device loop, and a coalesced batch —
    on a tiny linear problem.  Raises without a device or the library: there is no CPU fallback."""
    from . import _lib
    from .dynamics import LTI

    lib = _lib.load()
    if lib.pmpc_device_count() < 1:
        raise _lib.PmpcError("pmpc_b200.remote.precompilation_call: no CUDA device (there is no CPU fallback)")
    rng = np.random.default_rng(0)
    p = Problem(N=20, xdim=4, udim=2, x0=rng.standard_normal(4), res_tol=-1.0, max_it=5, verbose=False)
    A = 0.3 * rng.standard_normal((p.xdim, p.xdim)) + 0.4 * np.eye(p.xdim)
    B = np.array([[0.0, 0.0, 0.0, 0.3], [0.0, 0.0, 0.3, 0.0]]).T
    p.u_l, p.u_u = -10 * np.ones((p.N, p.udim)), 10 * np.ones((p.N, p.udim))
    p.solver_settings = dict(p.solver_settings, solver="osqp")

    def host_fn(x, u, p_=None):
        return ((A @ x[..., None])[..., 0] + (B @ u[..., None])[..., 0],
                np.broadcast_to(A, x.shape[:-1] + A.shape), np.broadcast_to(B, x.shape[:-1] + B.shape))

    for fn in (host_fn, LTI(A, B)):
        p.f_fx_fu_fn = fn
        scp_mpc.solve(**p)
    scp_mpc.solve_problems([dict(p), dict(p)])


# -- redis registry (one host: workers are reachable on the loopback interface only) ---------------
def get_redis_connection(redis_config: Optional[Dict[str, Any]] = None):
    return _redis_connect(redis_config)


def _redis_key(port: int) -> str:
    return "pmpc_worker_%s_%d/%s:%d" % (HOSTNAME, os.getpid(), BIND_HOST, port)


def set_redis_status(rconn, port: int) -> None:
    """Key with a 60 s time to live, refreshed by the serving loop every 10 s (``remote.py:186-194``)."""
    if redis is not None and rconn is not None:
        try:
            rconn.set(_redis_key(port), "%s:%d" % (BIND_HOST, port))
            rconn.expire(_redis_key(port), 60)
        except redis.ConnectionError:
            pass


def unset_redis_status(rconn, port: int) -> None:
    if redis is not None and rconn is not None:
        try:
            rconn.delete(_redis_key(port))
        except redis.ConnectionError:
            pass


# -- worker process --------------------------------------------------------------------------------
def _server(exit_flag, status_flag, port=DEFAULT_PORT, redis_config: Optional[Dict[str, Any]] = None,
            options: Optional[Dict[str, Any]] = None):
    """Body of the worker process (``remote.py:209-283``).  ``status_flag`` is the heartbeat the supervisor reads: the
    time of the last loop turn, a time in the far past for "could not start".  ``options``: ``device`` (the GPU this
    worker owns), ``slots``, ``batch_window``, ``precompile``, ``verbose``."""
    signal.signal(signal.SIGINT, signal.SIG_IGN)
    opts = dict(options or {})
    if opts.get("device") is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(opts["device"])  # before the library creates its context
    worker = Worker(port, slots=opts.get("slots", 1), batch_window=opts.get("batch_window", 1e-3),
                    verbose=opts.get("verbose", False))
    try:
        worker.bind()
    except zmq.ZMQError:
        print("Could not bind to port(s) %s" % worker.ports)
        status_flag.value = time.time() - 1e5
        return
    if opts.get("precompile", True):
        status_flag.value = time.time() + 60.0  # a minute of grace: a fresh box pages the CUDA libraries in
        try:
            precompilation_call()
        except Exception:
            print(traceback.format_exc())
            status_flag.value = time.time() - 1e5
            worker.close()
            return
    rconn, redis_update = get_redis_connection(redis_config), time.time() - 10.0
    while not exit_flag.value:
        status_flag.value = time.time()
        if time.time() - redis_update > 10.0:
            for p in worker.ports:
                set_redis_status(rconn, p)
            redis_update = time.time()
        if worker.serve_once(100):
            gc.collect()
    for p in worker.ports:
        unset_redis_status(rconn, p)
    worker.close()


class Server:
    """Handle of one worker process (``remote.py:286-318``): ``start`` / ``stop`` / ``is_alive`` / ``kill``.  The
    process is spawned, not forked: a parent that has touched CUDA cannot fork a child that uses it."""

    def __init__(self, port=DEFAULT_PORT, redis_config: Optional[Dict[str, Any]] = None, **options):
        self.port, self.redis_config, self.options = port, redis_config, options
        self.process = self.exit_flag = self.status_flag = None
        self.start()

    def start(self):
        import multiprocessing

        ctx = multiprocessing.get_context("spawn")
        self.exit_flag, self.status_flag = ctx.Value("b", False), ctx.Value("d", time.time())
        self.process = ctx.Process(target=_server, args=(self.exit_flag, self.status_flag, self.port, self.redis_config,
                                                         self.options), daemon=True)
        try:
            self.old_signal_handler = signal.signal(signal.SIGINT, self.sighandler)
        except ValueError:  # not the main thread
            self.old_signal_handler = None
        self.process.start()

    def stop(self, timeout: float = 30.0):
        """Ask the serving loop to leave and wait for the process; one that does not leave within ``timeout`` (stuck in
        a solve) is killed, so that stopping never hangs the caller."""
        if self.process is not None:
            self.exit_flag.value = True
            self.process.join(timeout)
            if self.process.is_alive():
                self.process.kill()
                self.process.join()
            self.process.close()
            self.process = self.exit_flag = self.status_flag = None

    def is_alive(self):
        return self.status_flag is not None and time.time() - self.status_flag.value < 60.0

    def failed(self):
        """True when the process reported that it could not start (port taken, or no device for the warm-up)."""
        return self.status_flag is not None and time.time() - self.status_flag.value > 1e4

    def wait_ready(self, timeout: float = 120.0) -> bool:
        """Block until the worker answers ``server_info`` (it binds before warming up, so a request sent earlier would
        simply wait).  False if it failed to start or ``timeout`` passed."""
        t0 = time.time()
        while time.time() - t0 < timeout and not self.failed() and self.process is not None and self.process.is_alive():
            poll = call("server_info", BIND_HOST, self.port, False)
            t1 = time.time()
            while time.time() - t1 < 2.0:
                ret = poll()
                if not (isinstance(ret, str) and ret == NOT_ARRIVED_YET):
                    return True
                time.sleep(0.01)
            poll.ctx.destroy()
        return False

    def kill(self):
        if self.process is not None:
            self.process.kill()
            rconn = get_redis_connection(self.redis_config)
            for i in range(max(1, int(self.options.get("slots", 1)))):
                unset_redis_status(rconn, self.port + i)

    def sighandler(self, sig, frame):
        self.stop()
        if callable(self.old_signal_handler):
            self.old_signal_handler(sig, frame)


SERVERS: Dict[int, Server] = dict()


def start_server(port: int = DEFAULT_PORT, verbose: bool = False, redis_config: Optional[Dict[str, Any]] = None,
                 **options):
    """Start a worker process on ``port`` (``remote.py:123-130``).  ``options``: ``device``, ``slots`` (the worker
    answers on ``port .. port+slots-1``), ``batch_window``, ``precompile``."""
    if port in SERVERS:
        raise RuntimeError("PMPC server on this port already exits")
    if verbose:
        print("Starting PMPC server on port: %d" % port)
    SERVERS[port] = Server(port, redis_config, verbose=verbose, **options)


def stop_servers():
    for port in list(SERVERS):
        SERVERS.pop(port).stop()


def main(argv=None):
    """``python -m pmpc_b200.remote``: the reference's worker launcher (``remote.py:456-517``) with GPUs in the place of
    cores.  ``--worker-num 0`` starts one worker per visible GPU (the reference: one per physical core); worker ``i``
    owns GPU ``i mod #GPUs`` and the ports ``port + i*slots ...``.  A worker whose heartbeat is older than 60 s is
    killed and, with ``--resurrect``, replaced by a new one on fresh ports."""
    from . import _lib

    parser = ArgumentParser()
    parser.add_argument("--port", "-p", type=int, default=DEFAULT_PORT, help="first TCP port (loopback interface)")
    parser.add_argument("--verbose", "-v", action="store_true")
    parser.add_argument("--worker-num", "-n", type=int, default=1, help="worker processes; 0 = one per visible GPU")
    parser.add_argument("--slots", type=int, default=16, help="ports per worker = jobs a stock client keeps in flight on it")
    parser.add_argument("--batch-window-ms", type=float, default=1.0)
    parser.add_argument("--resurrect", type=bool, default=True, help="Whether to resurrect dead workers.")
    parser.add_argument("--redis-host", type=str, default=None)
    parser.add_argument("--redis-port", type=str, default=None)
    parser.add_argument("--redis-password", type=str, default=None)
    args = parser.parse_args(argv)
    n_gpu = _lib.device_count()
    if n_gpu < 1:
        raise _lib.PmpcError("pmpc_b200.remote: no CUDA device; a worker cannot serve anything (there is no CPU fallback)")
    if args.worker_num == 0:
        args.worker_num = n_gpu
    redis_config = copy(REDIS_CONFIG)
    for key in ("host", "port", "password"):
        if getattr(args, "redis_" + key) is not None:
            redis_config[key] = getattr(args, "redis_" + key)
    # Ctrl-C and `kill` both end in the orderly shutdown below, also when the launcher was started with SIGINT ignored
    # (a background job of a non-interactive shell)
    def interrupt(sig, frame):
        raise KeyboardInterrupt

    signal.signal(signal.SIGINT, signal.default_int_handler)
    signal.signal(signal.SIGTERM, interrupt)
    try:
        for i in range(args.worker_num):
            start_server(args.port + i * args.slots, verbose=args.verbose, redis_config=redis_config, device=i % n_gpu,
                         slots=args.slots, batch_window=1e-3 * args.batch_window_ms)
        while True:
            for port, server in list(SERVERS.items()):
                if not server.is_alive():
                    print("Killing server on port %d" % port)
                    server.kill()
                    del SERVERS[port]
                    if args.resurrect:
                        new_port = max(list(SERVERS.keys()) + [port]) + args.slots
                        print("Restarting dead server on port %d" % new_port)
                        options = {k: v for k, v in server.options.items() if k != "verbose"}
                        start_server(new_port, verbose=args.verbose, redis_config=redis_config, **options)
            time.sleep(1.0)
    except KeyboardInterrupt:
        pass
    finally:
        signal.signal(signal.SIGINT, signal.SIG_IGN)  # a second Ctrl-C must not cut the shutdown short
        for port in sorted(SERVERS):
            print("Stopping server on port %d" % port, flush=True)
            SERVERS.pop(port).stop()


if __name__ == "__main__":
    main()

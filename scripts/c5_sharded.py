"""BASELINE config 5 as specified: ONE consensus problem (n=32, m=8, N=200, M=1024, Nc=5, slew, u box) with its
particles sharded over the GPUs of the node (torchrun, one process per GPU); the consensus terms are exchanged inside
the SCP kernel over NVLink.  Prints one JSON line on rank 0.   torchrun --nproc-per-node 8 scripts/c5_sharded.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from fixtures import config_c5
from pmpc_b200 import parallel

M = int(os.environ.get("C5_M", 1024)); N = int(os.environ.get("C5_N", 200))
rank, world, local = parallel.env_rank_world()
dist = parallel.init_process_group("gloo")
d, A, Bm = config_c5(M=M, N=N)
Nc = d["solver_settings"]["Nc"]
kw = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"], reg_x=d["reg_x"],
          reg_u=d["reg_u"], slew_rate=d["slew_rate"], max_it=d["max_it"], res_tol=d["res_tol"], Nc=Nc, ltv=(A, Bm, None))
reps = int(os.environ.get("C5_REPS", 3))
kms = []
for rep in range(reps):  # the first solve pays for CUDA's lazy module load on every rank (the kernels wait for each other)
    t = time.time()
    X, U, info = parallel.solve_consensus_sharded(None, d["Q"], d["R"], d["x0"], dist=dist, device=local, **kw)
    wall = time.time() - t
    kms.append(info["kernel_ms"])
info["kernel_ms"] = min(kms[1:]) if reps > 1 else kms[0]
if rank == 0:
    f = np.einsum("mjrt,mjt->mjr", A, X[:, :-1]) + np.einsum("mjrt,mjt->mjr", Bm, U)
    print(json.dumps(dict(config="C5 LTV n=32 m=8 N=%d M=%d Nc=%d sharded over %d GPUs" % (N, M, Nc, world), kernel_ms=info["kernel_ms"], kernel_ms_all=kms,
                          wall_s=wall, scp_iters=info["scp_iters"], status=info["status"], passes=info["polish_passes"],
                          dyn_violation=float(np.abs(f - X[:, 1:]).max()), consensus=bool(np.all(U[:, :Nc] == U[:1, :Nc])),
                          u_in_box=bool(np.abs(U).max() <= 1 + 1e-9))), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()

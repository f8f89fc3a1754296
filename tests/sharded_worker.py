"""Worker of the multi-GPU particle-sharding test (launched by torchrun, one process per GPU).

Solves ONE consensus problem with its particles sharded over the ranks (in-kernel exchange of the
consensus terms over peer memory) and, on rank 0, the same problem unsharded on a single GPU; writes
the comparison as JSON.  usage: sharded_worker.py out.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fixtures import config_c2  # noqa: E402

import pmpc_b200  # noqa: E402
from pmpc_b200 import parallel  # noqa: E402


def main():
    out = sys.argv[1]
    rank, world, local = parallel.env_rank_world()
    dist = parallel.init_process_group("gloo")
    res = {}
    # "wide": control box +-5, so the shared controls are interior and the exchanged sums decide them
    for name, (M, N, Nc, box) in dict(c2_small=(64, 20, 1, 1.0), c2_full=(128, 50, 1, 1.0), nc3=(32, 16, 3, 1.0),
                                      wide=(64, 20, 2, 5.0)).items():
        d, p = config_c2(M, N)
        d["u_l"], d["u_u"] = box * d["u_l"], box * d["u_u"]
        kw = dict(X_ref=d["X_ref"], U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"], u_u=d["u_u"],
                  reg_x=d["reg_x"], reg_u=d["reg_u"], max_it=d["max_it"], res_tol=d["res_tol"], Nc=Nc)
        t = time.time()
        X, U, info = parallel.solve_consensus_sharded(p, d["Q"], d["R"], d["x0"], dist=dist, device=local, **kw)
        wall = time.time() - t
        if rank == 0:
            ss = dict(d["solver_settings"], Nc=Nc)
            X1, U1, data1 = pmpc_b200.solve(pmpc_b200.DubinsCar(p), d["Q"], d["R"], d["x0"], X_ref=d["X_ref"],
                                            U_ref=d["U_ref"], X_prev=d["X_prev"], U_prev=d["U_prev"], u_l=d["u_l"],
                                            u_u=d["u_u"], reg_x=d["reg_x"], reg_u=d["reg_u"], max_it=d["max_it"],
                                            res_tol=d["res_tol"], solver_settings=ss)
            res[name] = dict(err_X=float(np.abs(X - X1).max()), err_U=float(np.abs(U - U1).max()),
                             scp_iters=info["scp_iters"], scp_iters_single=len(data1["hist"]), status=info["status"],
                             consensus=bool(np.all(U[:, :Nc] == U[:1, :Nc])), kernel_ms=info["kernel_ms"],
                             kernel_ms_single=data1["device"]["kernel_ms"], wall_s=wall, world=world)
    if rank == 0:
        with open(out, "w") as fh:
            json.dump(res, fh, indent=1)
        print(json.dumps(res))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

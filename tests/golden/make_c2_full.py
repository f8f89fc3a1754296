"""Generates tests/golden/scp_c2_full.npz: the converged solution of BASELINE config C2 at FULL size (Dubins N=50,
M=128, Nc=1, seed 1, res_tol=1e-5, max_it=100) from the CPU oracle — the reference's SCP loop (oracle.scp follows
pmpc/scp_mpc.py:297-442) around the exact QP oracle (oracle.lqp: lqp_utils.jl:2-393 assembly + KKT-certified solve).
The GPU test `test_full_size_c2_matches_exact_oracle_golden` asserts X, U to 1e-5 relative, the objective to 1e-6
relative and the iteration count against it.  About 2 minutes on one core:

    python tests/golden/make_c2_full.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from fixtures import config_c2
from oracle import dubins, scp as oscp

d, p = config_c2()
ss = d.pop("solver_settings")
t = time.time()
X, U, data = oscp.scp_solve(dubins.make_f_fx_fu_fn(p), d.pop("Q"), d.pop("R"), d.pop("x0"), solver_settings=ss, qp="exact", **d)
hist = np.array([[h["obj"], h["resid"]] for h in data["hist"]])
np.savez_compressed(os.path.join(HERE, "scp_c2_full.npz"), X=X, U=U, hist=hist, p=p, seconds=time.time() - t)
print("iters", len(hist), "obj", hist[-1, 0], "resid", hist[-1, 1], "seconds", time.time() - t)

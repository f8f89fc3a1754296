"""Generates tests/golden/c2_reference_iters.json: how many SCP iterations the CPU restatement of the
reference needs on BASELINE config C2 (Dubins N=50 M=128 Nc=1, seed 1) to reach res_tol=1e-5, with
(a) the reference's solver restated at its default settings (oracle.osqp_port, eps=1e-3) and
(b) the exact QP oracle (oracle.lqp), (c) the plain-C restatement (oracle/c/pmpc_oracle.c, OSQP 0.6.2 defaults).
bench.py scales its bounded CPU samples with (c).
Takes ~2 x 110 s + 40 s on one core:  python tests/golden/make_c2_iters.py"""
import json, os, sys, time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from fixtures import config_c2
from oracle import dubins, scp as oscp

out = {}
for key, qp in (("osqp_port", "osqp"), ("exact", "exact")):
    d, p = config_c2()
    ss = d.pop("solver_settings")
    t = time.time()
    X, U, data = oscp.scp_solve(dubins.make_f_fx_fu_fn(p), d.pop("Q"), d.pop("R"), d.pop("x0"), solver_settings=ss, qp=qp, **d)
    out[key + "_scp_iters"] = len(data["hist"])
    out[key + "_seconds_1core"] = round(time.time() - t, 1)
    out[key + "_final_resid"] = float(data["hist"][-1]["resid"])
    out[key + "_final_obj"] = float(data["hist"][-1]["obj"])
from oracle import cport

d, p = config_c2()
kw = {k: d[k] for k in ("X_ref", "U_ref", "X_prev", "U_prev", "u_l", "u_u", "reg_x", "reg_u", "max_it", "res_tol")}
t = time.time()
X, U, h = cport.scp_dubins(p, d["Q"], d["R"], d["x0"], Nc=1, **kw)
out.update(c_port_scp_iters=len(h), c_port_seconds_1core=round(time.time() - t, 1), c_port_final_resid=float(h[-1, 1]),
           c_port_final_obj=float(h[-1, 0]),
           c_port_note="oracle/c/pmpc_oracle.c (OSQP 0.6.2 restated at its defaults), one core of the build container")
json.dump(out, open(os.path.join(HERE, "c2_reference_iters.json"), "w"), indent=1)
print(out)

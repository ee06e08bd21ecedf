"""Developer probe (gpurun, 1 GPU): matvec counts of the device thick-restart
Lanczos against the reference's eigensolver (scipy ARPACK, ncv=4, tol=1e-3,
quimb/linalg/scipy_linalg.py:113-128) on THE SAME local problems of a first
right sweep from a random state, L = 100, chi = 1024."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quimb_b200 as qb
from quimb_b200.contract import contract_pair
from quimb_b200.dmrg import EffHam2
from quimb_b200.lanczos import eigh_arpack_host_driver, eigh_lanczos
from quimb_b200.mps import env_left_step

L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
probe_sites = [int(s) for s in (sys.argv[3].split(",") if len(sys.argv) > 3 else "12,14,16".split(","))]
d = qb.DMRG2(qb.mpo_ham_heis(L), chi, cutoffs=0.0, mpo_shape="lrdu", seed=2)
d.right_canonize()
d._init_right_envs()
d.lenv = {0: d._ones_env()}
rows = []
for i in range(max(probe_sites) + 1):
    if i > 0:
        d.lenv[i] = env_left_step(d.lenv[i - 1], d._k[i - 1], d.ham[i - 1])
        d.lenv.pop(i - 1, None)
    if i in probe_sites:
        A, B = d._k[i], d._k[i + 1]
        dims = (A.shape[0], A.shape[1], B.shape[1], B.shape[2])
        v0 = qb.Array(contract_pair(A.t, [0, 1, 9], B.t, [9, 2, 3], [0, 1, 2, 3]))
        row = {"site": i, "dims": dims}
        for ncv in (4, 8, 16, 32):
            H = EffHam2(d.lenv[i], d.ham[i], d.ham[i + 1], d.renv[i + 1], dims)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            th, x, info = eigh_lanczos(H, v0, which="SA", ncv=ncv, tol=1e-3, return_info=True,
                                       min_steps=4)
            torch.cuda.synchronize()
            row[f"device_ncv{ncv}"] = {"matvecs": H.nmatvec, "theta": th, "resid": info["resid"],
                                       "restarts": info["restarts"], "s": time.perf_counter() - t0}
        for ncv in (4, 8):
            H = EffHam2(d.lenv[i], d.ham[i], d.ham[i + 1], d.renv[i + 1], dims)
            t0 = time.perf_counter()
            th, x, info = eigh_arpack_host_driver(H, v0, which="SA", ncv=ncv, tol=1e-3)
            row[f"arpack_ncv{ncv}"] = {"matvecs": H.nmatvec, "theta": th,
                                       "s": time.perf_counter() - t0}
        rows.append(row)
        print(json.dumps(row), flush=True)
    d._update_local_state(i, "right", max_bond=chi, cutoff=0.0, cutoff_mode="sum2", method="svd")
    d.renv.pop(i + 1, None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"L": L, "chi": chi, "rows": rows, "sweep_nmatvecs": d.nmatvecs},
          open("gpurun_out/lanczos_compare.json", "w"), indent=1)

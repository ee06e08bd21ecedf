"""Developer probe (gpurun): matvec count / time of DMRG2 sweeps vs device Lanczos basis size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, quimb_b200 as qb
from oracle import dmrg_np as dm
L, chi = int(sys.argv[1]), int(sys.argv[2])
mpo = dm.mpo_heis(L)
for ncv in (4, 6, 8, 12, 16):
    d = qb.DMRG2(mpo, chi, cutoffs=0.0, mpo_shape="lrdu", seed=2)
    d.opts["device_eig_ncv"] = ncv
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e1 = d.sweep_right(canonize=True, max_bond=chi, cutoff=0.0)
    torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    mv1 = sum(d.nmatvecs)
    e2 = d.sweep_right(canonize=True, max_bond=chi, cutoff=0.0)
    torch.cuda.synchronize(); t2 = time.perf_counter() - t0 - t1
    print(f"ncv={ncv:2d} sweep1 {t1:7.2f}s matvecs {mv1:6d} E={e1:.8f} | sweep2 {t2:7.2f}s matvecs {sum(d.nmatvecs) - mv1:6d} E={e2:.8f}", flush=True)

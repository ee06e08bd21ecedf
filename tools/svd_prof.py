"""SVD profile helper (GPU box): sweeps, time, and -- with QB_TRACE=1 -- the
phase breakdown of jacobi_pair_kernel from its in-kernel %globaltimer stamps
(cluster 0 of each round of the last sweep).

phases: 0 entry, 1 Gram streamed, 2 Gram reduced over the cluster + off-norm,
        3 32x32 eigen-solve done, 4 sorted, 5 rotation applied, 6 cluster exit
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quimb_b200 as qb
from quimb_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
X = qb.Array(torch.randn(n, n, dtype=torch.float64, device='cuda'))
qb.linalg.svd(X)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
U, s, VH, sw = qb.linalg.svd(X, return_sweeps=True)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"n": n, "sweeps": sw, "ms": e0.elapsed_time(e1)}))
if os.environ.get("QB_TRACE"):
    cnt = 17 * 8192
    buf = (ctypes.c_ulonglong * cnt)()
    assert _lib.load().qb_debug_trace_read(buf, cnt) == 0
    t = np.frombuffer(buf, dtype=np.uint64)[16 * 8192:16 * 8192 + 1024 * 8].reshape(1024, 8).astype(np.int64)
    ok = (t[:, 0] > 0) & (t[:, 6] > t[:, 0]) & (t[:, 3] > t[:, 2])   # rounds that rotated
    names = ["gram_stream", "reduce+offnorm", "eigensolve", "sort", "apply", "cluster_exit"]
    out = {}
    for ph in range(6):
        a, b = t[ok, ph], t[ok, ph + 1]
        m = (b >= a) & (a > 0)
        if m.any():
            out[names[ph]] = round(float(np.median((b - a)[m])) / 1e3, 2)
    starts = np.sort(t[ok, 0])
    out["round_period_us"] = round(float(np.median(np.diff(starts))) / 1e3, 2) if len(starts) > 2 else None
    out["rounds_traced"] = int(ok.sum())
    print(json.dumps(out))

"""QR timing / accuracy at the shapes of the hot path (GPU box):
DMRG canonisation (2048 x 1024), the SVD preconditioner (2048^2) and the tall
boundary tensors of BASELINE configs[4] (16384 x 2048, 16384 x 1024)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quimb_b200 as qb

g = torch.Generator(device="cuda").manual_seed(2)
rows = []
for m, n, check in [(2048, 1024, True), (2048, 2048, False), (4096, 512, True), (8192, 1024, False),
                    (16384, 1024, False), (16384, 2048, True), (100, 37, True), (5000, 130, True)]:
    X = qb.Array(torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g))
    qb.linalg.qr(X, stabilized=True); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); Q, R = qb.linalg.qr(X, stabilized=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    row = {"m": m, "n": n, "ms": min(ts), "tflops_R_plus_Q": 4.0 * m * n * n / (min(ts) * 1e-3) / 1e12}
    if check:
        q, r, x = Q.to_numpy(), R.to_numpy(), X.to_numpy()
        row["orth"] = float(np.abs(q.T @ q - np.eye(n)).max())
        row["recon"] = float(np.abs(q @ r - x).max())
        row["tril"] = float(np.abs(np.tril(r, -1)).max())
        row["diag_min"] = float(np.diag(r).min())
    rows.append(row)
    print(json.dumps(row), flush=True)

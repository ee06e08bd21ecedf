"""Developer probe (gpurun): time the pieces of one DMRG2 two-site update at
chi=1024 and the QR / SVD kernels at DMRG sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quimb_b200 as qb
from quimb_b200.dmrg import EffHam2

def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d, w = 2, 5
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: qb.Array(torch.randn(*s, dtype=torch.float64, device="cuda", generator=g))
L, R = rn(chi, w, chi), rn(chi, w, chi)
W1, W2 = rn(w, w, d, d), rn(w, w, d, d)
x = rn(chi * d * d * chi)
H = EffHam2(L, W1, W2, R, (chi, d, d, chi))
print("matvec ms", t_ms(lambda: H.matvec(x)), "GF", H.flops() / 1e9, flush=True)
A = rn(chi, d, chi)
print("env_left ms", t_ms(lambda: qb.env_left_step(L, A, W1)), flush=True)
X = rn(chi * d, d * chi)
for m, n in [(chi * d, chi), (chi * d, chi * d)]:
    M = rn(m, n)
    print(f"qr {m}x{n} ms", t_ms(lambda: qb.linalg.qr(M, stabilized=True), 2), flush=True)
t0 = time.perf_counter()
U, s, VH, sweeps = qb.linalg.svd(X, return_sweeps=True); torch.cuda.synchronize()
print(f"svd {chi*d}x{chi*d} ms", (time.perf_counter() - t0) * 1e3, "sweeps", sweeps, flush=True)
t0 = time.perf_counter()
U, s, VH, sweeps = qb.linalg.svd(X, return_sweeps=True); torch.cuda.synchronize()
print(f"svd again ms", (time.perf_counter() - t0) * 1e3, "sweeps", sweeps, flush=True)
xs = X.to_numpy()
t0 = time.perf_counter(); sref = np.linalg.svd(xs, compute_uv=False); print("numpy svd (vals only) s", time.perf_counter() - t0)
print("sv err", np.max(np.abs(s.to_numpy() - sref)) / sref[0])
t0 = time.perf_counter(); np.linalg.svd(xs, full_matrices=False); print("numpy svd full s", time.perf_counter() - t0)
t0 = time.perf_counter(); np.linalg.qr(xs[:, :chi]); print("numpy qr s", time.perf_counter() - t0)
# lanczos on the effective hamiltonian (symmetrised random): count matvecs
Ls = qb.Array(L.t + L.t.permute(2, 1, 0)); Rs = qb.Array(R.t + R.t.permute(2, 1, 0))
W1s = qb.Array(W1.t + W1.t.permute(0, 1, 3, 2)); W2s = qb.Array(W2.t + W2.t.permute(0, 1, 3, 2))
Hs = EffHam2(Ls, W1s, W2s, Rs, (chi, d, d, chi))
t0 = time.perf_counter()
th, v, info = qb.eigh_lanczos(Hs, x, ncv=4, tol=1e-3, return_info=True); torch.cuda.synchronize()
print("lanczos ms", (time.perf_counter() - t0) * 1e3, info, flush=True)

"""Developer check (gpurun): tcgen05 Ozaki engine vs the DMMA engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quimb_b200.contract import contract_pair
dev = "cuda"
def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
g = torch.Generator(device=dev).manual_seed(0)
def check(name, a, la, b, lb, lc):
    ref = contract_pair(a, la, b, lb, lc, engine=1)
    out = contract_pair(a, la, b, lb, lc, engine=2)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item(); sc = ref.abs().max().item()
    print(f"{name:28s} rel err {err / sc:.3e}  (abs {err:.3e}, scale {sc:.3e})", flush=True)
    return err / sc
for (m, k, n) in [(128, 128, 64), (256, 256, 128), (384, 512, 320), (1000, 777, 333), (1024, 1024, 2048)]:
    a = torch.randn(m, k, dtype=torch.float64, device=dev, generator=g)
    b = torch.randn(k, n, dtype=torch.float64, device=dev, generator=g)
    check(f"gemm {m}x{k}x{n}", a, [0, 1], b, [1, 2], [0, 2])
a = torch.randn(300, 200, dtype=torch.float64, device=dev, generator=g)
b = torch.randn(200, 150, dtype=torch.float64, device=dev, generator=g)
check("gemm TN", a.t().contiguous().t(), [0, 1], b, [1, 2], [0, 2])
A = torch.randn(32, 20, 24, 28, dtype=torch.float64, device=dev, generator=g)
B = torch.randn(28, 16, 20, 18, dtype=torch.float64, device=dev, generator=g)
check("perm acbd,dfce->abef", A, [0, 2, 1, 3], B, [3, 5, 2, 4], [0, 1, 4, 5])
# badly scaled rows / columns
a = torch.randn(512, 512, dtype=torch.float64, device=dev, generator=g) * torch.logspace(-8, 8, 512, dtype=torch.float64, device=dev)[:, None]
b = torch.randn(512, 512, dtype=torch.float64, device=dev, generator=g) * torch.logspace(-6, 6, 512, dtype=torch.float64, device=dev)[None, :]
ref = contract_pair(a, [0, 1], b, [1, 2], [0, 2], engine=1); out = contract_pair(a, [0, 1], b, [1, 2], [0, 2], engine=2)
print("scaled rows/cols: max rel (elementwise vs row*col scale)", ((out - ref).abs() / (a.abs().max(1, keepdim=True).values * b.abs().max(0, keepdim=True).values * 512 ** 0.5)).max().item(), flush=True)
chi, d, w = 1024, 2, 5
for name, (m, k, n) in {"mps step (1024,1024,2048)": (1024, 1024, 2048), "dmrg L.x (5120,1024,4096)": (5120, 1024, 4096), "4096^3": (4096, 4096, 4096)}.items():
    a = torch.randn(m, k, dtype=torch.float64, device=dev, generator=g)
    b = torch.randn(k, n, dtype=torch.float64, device=dev, generator=g)
    c = torch.empty(m, n, dtype=torch.float64, device=dev)
    fl = 2 * m * n * k
    t1 = t_ms(lambda: contract_pair(a, [0, 1], b, [1, 2], [0, 2], out=c, engine=1))
    t2 = t_ms(lambda: contract_pair(a, [0, 1], b, [1, 2], [0, 2], out=c, engine=2))
    print(f"{name}: dmma {t1:.3f} ms {fl / t1 / 1e9:.1f} TF/s | ozaki {t2:.3f} ms {fl / t2 / 1e9:.1f} TF/s", flush=True)

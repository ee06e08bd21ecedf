import sys; sys.path.insert(0, '/root/repo')
import torch, quimb_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
X = qb.Array(torch.randn(n, n, dtype=torch.float64, device='cuda'))
U, s, VH, sw = qb.linalg.svd(X, return_sweeps=True)
torch.cuda.synchronize(); print("sweeps", sw)

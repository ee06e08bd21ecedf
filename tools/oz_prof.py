import sys; sys.path.insert(0, '/root/repo')
import torch
from quimb_b200.contract import contract_pair
a = torch.randn(4096, 4096, dtype=torch.float64, device='cuda'); b = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
c = torch.empty(4096, 4096, dtype=torch.float64, device='cuda')
for _ in range(3): contract_pair(a, [0, 1], b, [1, 2], [0, 2], out=c, engine=2)
torch.cuda.synchronize()

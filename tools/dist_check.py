"""Run under torchrun on N GPUs: slice-parallel contraction of a random
circuit amplitude (BASELINE configs[3]-like) with ONE NCCL all-reduce of the
scalar result; rank 0 checks it against the dense state-vector value."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import quimb_b200 as qb
from tests.circuit_util import random_circuit_amplitude

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nq, depth = 14, 10
arrays, inputs, output, amp = random_circuit_amplitude(nq, depth, seed=11)
counts = {}
for t in inputs:
    for ix in t:
        counts[ix] = counts.get(ix, 0) + 1
inner = [ix for ix, c in counts.items() if c == 2]
sliced = inner[len(inner) // 2: len(inner) // 2 + 4]          # 16 independent slices
dev = [qb.asarray(a) for a in arrays]
torch.cuda.synchronize()
t0 = time.perf_counter()
total, mine = qb.dist.contract_sliced(dev, inputs, output, sliced, optimize="greedy")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
val = complex(total.item())
if rank == 0:
    ok = abs(val - amp) <= 1e-10
    print(json.dumps({"world": world, "slices": 16, "mine": len(mine), "amp": [val.real, val.imag],
                      "exact": [amp.real, amp.imag], "abs_err": abs(val - amp), "ok": bool(ok),
                      "seconds": dt}), flush=True)
    assert ok
if world > 1:
    dist.destroy_process_group()

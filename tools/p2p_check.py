"""Run under torchrun: checks the fused peer-memory exchange kernels
(csrc/p2p.cu) against torch.distributed on the same data and times both.

  N GPUs (NCCL):   torchrun --nproc-per-node N tools/p2p_check.py
  one GPU, 2 procs: torchrun --nproc-per-node 2 tools/p2p_check.py --one-gpu
                    (gloo group, both ranks on cuda:0: CUDA IPC between two
                    processes of the same device; kernels are time-sliced)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from quimb_b200.dist import BondShard

one_gpu = "--one-gpu" in sys.argv
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", 0 if one_gpu else lr)
torch.cuda.set_device(dev)
if one_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
out = {"world": world, "one_gpu": one_gpu}
n, cols = 1024, 4096
if one_gpu:
    n, cols = 256, 512
ref = BondShard(exchange="nccl")
px = BondShard(exchange="p2p")
lo, hi = px.slab(n)
g = torch.Generator(device=dev).manual_seed(5)
full = torch.randn((n, cols), dtype=torch.float64, device=dev, generator=g)
ok = True
for it in range(5):
    local = (full[lo:hi] * (it + 1)).contiguous()
    a = px.all_gather_rows(local, n, transient=True)
    b = ref.all_gather_rows(local, n)
    same = bool(torch.equal(a, b)) and bool(torch.equal(a, full * (it + 1)))
    ok &= same
    h = torch.arange(1, 25, dtype=torch.float64, device=dev) * (rank + 1) * 0.1 * (it + 1)
    h2 = h.clone()
    px.all_reduce_(h)
    ref.all_reduce_(h2)
    ok &= bool(torch.allclose(h, h2, rtol=1e-15, atol=0))
px.check()
out["exchange"] = px.exchange_name
out["equal"] = bool(ok)


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


local = full[lo:hi].contiguous()
reps = 5 if one_gpu else 50
out["allgather_ms"] = {"p2p": timeit(lambda: px.all_gather_rows(local, n, transient=True), reps),
                       "nccl": timeit(lambda: ref.all_gather_rows(local, n), reps)}
h = torch.ones(24, dtype=torch.float64, device=dev)
out["allreduce24_ms"] = {"p2p": timeit(lambda: px.all_reduce_(h), reps),
                         "nccl": timeit(lambda: ref.all_reduce_(h), reps)}
out["bytes"] = n * cols * 8
px.check()
if rank == 0:
    print(json.dumps(out), flush=True)
    assert ok
px.close()
dist.barrier()
dist.destroy_process_group()

#!/bin/bash
# round-2 session 12 (8 GPUs): parity subset on rank-0's GPU, then the N=8 bench line
tag=r02s12
out=gpurun_out
mkdir -p $out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -x -q -k "dmrg or lanczos or linop or size_parity or tebd" > $out/${tag}_pytest_subset.log 2>&1
tail -4 $out/${tag}_pytest_subset.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29771 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err
grep '^{' $out/${tag}_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['exchange'], d['shard_unit'], d['e2e'])"
tail -3 $out/${tag}_bench_8gpu.err | cut -c1-300

#!/bin/bash
# Multi-GPU evidence in one call.  Usage (N = 2, 4 or 8):
#   gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_session_multi.sh r02 8'
tag=${1:-r02}
n=${2:-8}
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
# 1. the bench contract at N GPUs (independent MPS replicas, weak scaling)
timeout 600 run bench.py --gpus $n --steps 5 --warmup 3 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
cat $out/${tag}_bench_${n}gpu.json
# 2. bond-sharded DMRG eigensolve (BASELINE configs[2]; NCCL all-gather per matvec)
timeout 900 run tools/bench_dmrg.py --L 30 --chi 1024 --shard > $out/${tag}_dmrg_shard_${n}gpu.log 2>&1
# 3. circuit amplitude, BASELINE configs[3] at full size: slices dealt to the ranks, one all-reduce
timeout 1500 run tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 64 \
    --out $out/${tag}_circuit_6x6_d24_${n}gpu.json > $out/${tag}_circuit_${n}gpu.log 2>&1
QB_ENGINE=stream timeout 1500 run tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 64 \
    --out $out/${tag}_circuit_6x6_d24_${n}gpu_stream.json >> $out/${tag}_circuit_${n}gpu.log 2>&1
# 4. two-sided PEPS boundary contraction (BASELINE configs[4]) on 2 ranks
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 \
    tools/bench_boundary.py --two-sided > $out/${tag}_boundary_two_sided.json 2> $out/${tag}_boundary_two_sided.err
ls -la $out | tail -12

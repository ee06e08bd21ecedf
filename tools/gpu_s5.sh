#!/bin/bash
# round-2 session 5 (8 GPUs): the N=8 bench line and BASELINE configs[3] at
# full size (6x6 qubits, depth 24: every slice, one all-reduce).
tag=r02s5
out=gpurun_out
mkdir -p $out
run() { local t=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run 400 bench.py --gpus 8 --steps 5 --warmup 3 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err
tail -c 2500 $out/${tag}_bench_8gpu.json; tail -4 $out/${tag}_bench_8gpu.err
run 300 bench.py --gpus 8 --steps 5 --warmup 3 --exchange nccl --no-e2e > $out/${tag}_bench_8gpu_nccl.json 2> $out/${tag}_bench_8gpu_nccl.err
tail -c 1200 $out/${tag}_bench_8gpu_nccl.json; tail -4 $out/${tag}_bench_8gpu_nccl.err
run 600 tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 32 --reps 0 --tree-file profiles/r02_tree_cfg4_w31.json \
    --out $out/${tag}_circuit_6x6_d24_8gpu.json > $out/${tag}_circuit_8gpu.log 2>&1
tail -3 $out/${tag}_circuit_8gpu.log | cut -c1-1200

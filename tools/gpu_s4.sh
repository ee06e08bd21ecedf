#!/bin/bash
# round-2 session 4 (2 GPUs): peer-memory exchange vs NCCL, the N=2 bench line
# (both exchanges), two-sided boundary contraction.
tag=r02s4
out=gpurun_out
mkdir -p $out
run() { local t=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run 200 tools/p2p_check.py > $out/${tag}_p2p_check.log 2>&1; tail -2 $out/${tag}_p2p_check.log | cut -c1-600
run 400 bench.py --gpus 2 --steps 5 --warmup 3 --exchange p2p > $out/${tag}_bench_2gpu_p2p.json 2> $out/${tag}_bench_2gpu_p2p.err
tail -c 2500 $out/${tag}_bench_2gpu_p2p.json; tail -4 $out/${tag}_bench_2gpu_p2p.err
run 400 bench.py --gpus 2 --steps 5 --warmup 3 --exchange nccl --no-e2e > $out/${tag}_bench_2gpu_nccl.json 2> $out/${tag}_bench_2gpu_nccl.err
tail -c 1500 $out/${tag}_bench_2gpu_nccl.json; tail -4 $out/${tag}_bench_2gpu_nccl.err
run 300 tools/bench_boundary.py --Lx 6 --Ly 6 --D 8 --chi 128 --reps 0 --two-sided > $out/${tag}_boundary_two_sided_6x6.json 2> $out/${tag}_boundary_two_sided.err
cut -c1-400 $out/${tag}_boundary_two_sided_6x6.json; tail -3 $out/${tag}_boundary_two_sided.err

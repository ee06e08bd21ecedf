#!/bin/bash
# round-2 session 9 (2 GPUs): BASELINE configs[4] at full size, two-sided
tag=r02s9
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 \
    tools/bench_boundary.py --Lx 10 --Ly 10 --D 8 --chi 256 --reps 0 --two-sided \
    > $out/${tag}_boundary_10x10_two_sided.json 2> $out/${tag}_boundary.err
grep '^{' $out/${tag}_boundary_10x10_two_sided.json | cut -c1-500; tail -3 $out/${tag}_boundary.err

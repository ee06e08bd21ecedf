#!/bin/bash
# round-2 session 11 (8 GPUs): BASELINE configs[2] as written -- one DMRG2 sweep,
# Heisenberg L=100 chi=1024, local eigensolves bond-sharded over the 8 ranks.
tag=r02s11n2
out=gpurun_out
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
    tools/bench_dmrg.py --L 100 --chi 1024 --shard > $out/${tag}_dmrg_shard2.log 2>&1
grep '^{' $out/${tag}_dmrg_shard2.log | tail -1 | cut -c1-900; tail -3 $out/${tag}_dmrg_shard2.log | cut -c1-300

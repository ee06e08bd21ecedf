#!/bin/bash
# round-2 session 1 (1 GPU): full parity tier (no -x), the new bench line, the
# peer-memory exchange between two processes of one GPU, cfg5 / cfg4 timings,
# Lanczos vs ARPACK matvec counts.
tag=r02s1
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
tail -15 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tools/p2p_check.py --one-gpu > $out/${tag}_p2p_onegpu.log 2>&1
tail -3 $out/${tag}_p2p_onegpu.log
bash tools/svd_tune.sh gpurun_out/${tag}_svd_tune.log > /dev/null 2>&1
grep -h '"ms"' gpurun_out/${tag}_svd_tune.log | cut -c1-300
timeout 400 python tools/lanczos_compare.py 100 1024 12,15 > $out/${tag}_lanczos_compare.log 2>&1
tail -3 $out/${tag}_lanczos_compare.log
timeout 400 python tools/bench_boundary.py --Lx 10 --Ly 10 --D 8 --chi 256 > $out/${tag}_boundary.json 2> $out/${tag}_boundary.err
tail -c 600 $out/${tag}_boundary.json; tail -3 $out/${tag}_boundary.err
timeout 300 python tools/bench_circuit.py --Lx 5 --Ly 5 --depth 16 --target-width 22 \
    --out $out/${tag}_circuit_5x5_d16.json > $out/${tag}_circuit.log 2>&1
QB_ENGINE=stream timeout 300 python tools/bench_circuit.py --Lx 5 --Ly 5 --depth 16 --target-width 22 \
    --out $out/${tag}_circuit_5x5_d16_stream.json >> $out/${tag}_circuit.log 2>&1
tail -6 $out/${tag}_circuit.log

#!/bin/bash
# One gpurun call that collects a round's single-GPU evidence.  Usage:
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r02'
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
# 1. parity tier (stops at the first failure like the driver)
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
tail -3 $out/${tag}_pytest_gpu.log
# 2. the bench line, both arms
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
cat $out/${tag}_bench.json
# 3. launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
    > $out/${tag}_bench_under_ncu.log 2>&1
# 4. callers either side of the path
timeout 900 python tools/bench_dmrg.py --L 30 --chi 1024 --no-cpu --left-sweep > $out/${tag}_dmrg_L30.log 2>&1
timeout 600 python tools/bench_boundary.py --Lx 10 --Ly 10 --D 8 --chi 256 > $out/${tag}_boundary.json 2> $out/${tag}_boundary.err
timeout 600 python tools/bench_circuit.py --Lx 5 --Ly 5 --depth 16 --target-width 22 \
    --out $out/${tag}_circuit_5x5_d16.json > $out/${tag}_circuit.log 2>&1
timeout 900 python tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 2 \
    --out $out/${tag}_circuit_6x6_d24_partial.json >> $out/${tag}_circuit.log 2>&1
ls -la $out | tail -20
# 5. streaming engine (opt-in) against the DMMA path on the same circuit tree
timeout 300 python -m pytest tests/test_gpu_zzz_stream.py -q > $out/${tag}_pytest_stream.log 2>&1; tail -2 $out/${tag}_pytest_stream.log
QB_ENGINE=stream timeout 600 python tools/bench_circuit.py --Lx 5 --Ly 5 --depth 16 --target-width 22 \
    --out $out/${tag}_circuit_5x5_d16_stream.json >> $out/${tag}_circuit.log 2>&1
QB_ENGINE=stream timeout 900 python tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 2 \
    --out $out/${tag}_circuit_6x6_d24_partial_stream.json >> $out/${tag}_circuit.log 2>&1
# 6. the MPO steps of the DMRG matvec are (w d = 10)-deep, 10-wide contractions over 2^21 rows:
#    same sweep with the streaming engine taking every eligible step
QB_ENGINE=stream timeout 900 python tools/bench_dmrg.py --L 30 --chi 1024 --no-cpu --left-sweep > $out/${tag}_dmrg_L30_stream.log 2>&1
tail -2 $out/${tag}_dmrg_L30.log $out/${tag}_dmrg_L30_stream.log

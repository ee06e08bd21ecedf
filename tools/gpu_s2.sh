#!/bin/bash
# round-2 session 2 (1 GPU): parity tier after the SVD rewrite, SVD configs,
# bench (e2e staging ring), boundary profile, 6x6 circuit slices.
tag=r02s2
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
tail -8 $out/${tag}_pytest_gpu.log
log=$out/${tag}_svd_tune.log; : > $log
run() { echo "== $*" >> $log; env "$@" timeout 100 python tools/svd_prof.py 2048 >> $log 2>&1; }
timeout 100 python tools/svd_prof.py 2048 --check >> $log 2>&1
run QB_JAC_GROUPS=4 QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_GROUPS=4 QB_JAC_CS=4 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_GROUPS=4 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_GROUPS=4 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=3
run QB_JAC_GROUPS=2 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_GROUPS=2 QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_GROUPS=1 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_GROUPS=4 QB_JAC_CS=2 QB_JAC_CH=64 QB_JAC_STG=4 QB_TRACE=1
run QB_JAC_GROUPS=4 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2 QB_TRACE=1
timeout 60 python tools/svd_prof.py 1024 --check >> $log 2>&1
grep -h '"ms"\|gram' $log | cut -c1-330
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s2_bench.json'))
print({k:d[k] for k in ('value','e2e','dmrg','shard_unit')})
PY
tail -3 $out/${tag}_bench.err
timeout 300 python tools/bench_boundary.py --Lx 6 --Ly 6 --D 4 --chi 64 --reps 0 --profile > $out/${tag}_boundary_6x6_D4_chi64_profile.json 2> $out/${tag}_boundary.err
timeout 300 python tools/bench_boundary.py --Lx 6 --Ly 6 --D 8 --chi 128 --reps 0 --profile > $out/${tag}_boundary_6x6_D8_chi128_profile.json 2>> $out/${tag}_boundary.err
cut -c1-1500 $out/${tag}_boundary_6x6_D8_chi128_profile.json; tail -3 $out/${tag}_boundary.err
timeout 400 python tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 1 --reps 0 \
    --out $out/${tag}_circuit_6x6_d24_1slice.json > $out/${tag}_circuit.log 2>&1
QB_ENGINE=stream timeout 400 python tools/bench_circuit.py --Lx 6 --Ly 6 --depth 24 --target-width 31 --max-slices 1 --reps 0 \
    --out $out/${tag}_circuit_6x6_d24_1slice_stream.json >> $out/${tag}_circuit.log 2>&1
tail -4 $out/${tag}_circuit.log | cut -c1-900

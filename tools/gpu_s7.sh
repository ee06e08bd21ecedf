#!/bin/bash
# round-2 session 7 (1 GPU): validation of the accumulation-free split path and the
# auto engine selection; bench line.
tag=r02s7
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
tail -8 $out/${tag}_pytest_gpu.log
log=$out/${tag}_svd.log; : > $log
timeout 120 python tools/svd_prof.py 2048 --check >> $log 2>&1
QB_SVD_ACCUMULATE=1 timeout 120 python tools/svd_prof.py 2048 >> $log 2>&1
timeout 60 python tools/svd_prof.py 1024 --check >> $log 2>&1
grep -h '"ms"' $log | cut -c1-500
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s7_bench.json'))
print({k:d[k] for k in ('value','e2e','dmrg','shard_unit')})
PY
tail -3 $out/${tag}_bench.err
timeout 300 python tools/bench_dmrg.py --L 100 --chi 1024 --no-cpu > $out/${tag}_dmrg_sweep_L100.log 2>&1; tail -1 $out/${tag}_dmrg_sweep_L100.log | cut -c1-700
timeout 300 python tools/bench_boundary.py --Lx 6 --Ly 6 --D 8 --chi 128 --reps 0 > $out/${tag}_boundary_6x6_D8_chi128.json 2> $out/${tag}_boundary.err
cut -c1-400 $out/${tag}_boundary_6x6_D8_chi128.json

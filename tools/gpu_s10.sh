#!/bin/bash
# round-2 session 10 (1 GPU): final validation -- the driver's own sequence
tag=r02s10
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
tail -5 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s10_bench.json')); r=json.loads([l for l in open('gpurun_out/r02s10_bench_ref.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['traffic']); print(d['dmrg']); print('ref', r['value'], r['cpu_baseline'])
PY
tail -3 $out/${tag}_bench.err

#!/bin/bash
# round-2 session 3 (1 GPU): blocked-WY QR validation + timing, Jacobi graph replay /
# stagger, matvec engine probe, boundary contraction at BASELINE configs[4] size.
tag=r02s3
out=gpurun_out
mkdir -p $out
timeout 200 python tools/qr_prof.py > $out/${tag}_qr_prof.log 2>&1; cat $out/${tag}_qr_prof.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_gpu.log 2>&1
tail -6 $out/${tag}_pytest_gpu.log
log=$out/${tag}_svd_tune.log; : > $log
run() { echo "== $*" >> $log; env "$@" timeout 100 python tools/svd_prof.py 2048 >> $log 2>&1; }
run QB_JAC_PROFILE=1 QB_JAC_GRAPH=0
run QB_JAC_PROFILE=1 QB_JAC_GRAPH=1
run QB_JAC_STAGGER=12000
run QB_JAC_STAGGER=20000
run QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_STAGGER=12000 QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_STAGGER=10000 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_PROFILE=1 QB_JAC_GRAPH=0 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
grep -h '"ms"\|qb jacobi' $log | cut -c1-260
timeout 200 python tools/matvec_probe.py > $out/${tag}_matvec_probe.log 2>&1; cat $out/${tag}_matvec_probe.log | cut -c1-300
timeout 300 python tools/bench_boundary.py --Lx 6 --Ly 6 --D 8 --chi 128 --reps 0 > $out/${tag}_boundary_6x6_D8_chi128.json 2> $out/${tag}_boundary.err
cut -c1-400 $out/${tag}_boundary_6x6_D8_chi128.json
timeout 900 python tools/bench_boundary.py --Lx 10 --Ly 10 --D 8 --chi 256 --reps 0 > $out/${tag}_boundary_10x10_D8_chi256.json 2>> $out/${tag}_boundary.err
cut -c1-400 $out/${tag}_boundary_10x10_D8_chi256.json; tail -3 $out/${tag}_boundary.err

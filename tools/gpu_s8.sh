#!/bin/bash
# round-2 session 8 (1 GPU): Jacobi configurations for the accumulation-free split
tag=r02s8
out=gpurun_out
mkdir -p $out
log=$out/${tag}_svd_tune.log; : > $log
run() { echo "== $*" >> $log; env "$@" timeout 100 python tools/svd_prof.py 2048 >> $log 2>&1; }
run QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_CS=4 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2
run QB_JAC_CS=2 QB_JAC_CH=64 QB_JAC_STG=4 QB_JAC_GROUPS=2
run QB_JAC_CS=2 QB_JAC_CH=32 QB_JAC_STG=3
run QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2 QB_JAC_STAGGER=12000
run QB_JAC_CS=2 QB_JAC_CH=64 QB_JAC_STG=4 QB_TRACE=1 QB_JAC_GRAPH=0
grep -h '"ms"\|gram' $log | cut -c1-420

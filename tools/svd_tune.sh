#!/bin/bash
# A/B of the Jacobi SVD configurations at 2048^2 (one process per configuration)
out=${1:-gpurun_out/svd_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python tools/svd_prof.py 2048 >> $out 2>&1; }
QB_JAC_MODE=v1 timeout 120 python tools/svd_prof.py 2048 --check >> $out 2>&1
timeout 120 python tools/svd_prof.py 2048 --check >> $out 2>&1
for g in 1 2 4; do for cs in 2 4 8; do
  run QB_JAC_GROUPS=$g QB_JAC_CS=$cs QB_JAC_CH=32 QB_JAC_STG=2
done; done
run QB_JAC_GROUPS=4 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=3
run QB_JAC_GROUPS=4 QB_JAC_CS=4 QB_JAC_CH=64 QB_JAC_STG=2
run QB_JAC_GROUPS=2 QB_JAC_CS=8 QB_JAC_CH=64 QB_JAC_STG=3
run QB_JAC_GROUPS=4 QB_JAC_CS=8 QB_JAC_CH=32 QB_JAC_STG=2 QB_TRACE=1
run QB_JAC_GROUPS=1 QB_JAC_CS=2 QB_JAC_CH=64 QB_JAC_STG=4 QB_TRACE=1
timeout 60 python tools/svd_prof.py 1024 --check >> $out 2>&1
timeout 60 python tools/svd_prof.py 512 --check >> $out 2>&1
timeout 60 python tools/svd_prof.py 100 --check >> $out 2>&1
cat $out

#!/bin/bash
# round-2 profiling session (1 GPU): ncu --set full captures of the dominant kernels
# and the launch list of the bench command; numbers under a profiler are never bench values.
tag=r02
out=gpurun_out
mkdir -p $out
timeout 200 python tools/oz_single_prof.py > $out/${tag}_oz_single_prof.log 2>&1; cat $out/${tag}_oz_single_prof.log | cut -c1-300
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:contract_f64_streamk -s 1000 -c 2 -f -o $out/prof_contract_${tag} \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-dmrg > $out/${tag}_ncu_contract.log 2>&1
timeout 300 $NCU -k regex:jacobi_round -s 700 -c 3 -f -o $out/prof_jacobi_${tag} \
    env QB_JAC_GRAPH=0 python tools/svd_prof.py 2048 > $out/${tag}_ncu_jacobi.log 2>&1
timeout 300 $NCU -k regex:ozaki_gemm -s 2 -c 2 -f -o $out/prof_ozaki_single_${tag} \
    python tools/oz_single_prof.py > $out/${tag}_ncu_ozaki.log 2>&1
timeout 300 $NCU -k regex:qr_panel -s 20 -c 2 -f -o $out/prof_qr_panel_${tag} \
    python tools/qr_prof.py > $out/${tag}_ncu_qr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-dmrg \
    > $out/${tag}_bench_under_ncu.log 2>&1
timeout 300 python tools/bench_dmrg.py --L 100 --chi 1024 --no-cpu > $out/${tag}_dmrg_sweep_L100.log 2>&1; tail -2 $out/${tag}_dmrg_sweep_L100.log | cut -c1-600
ls -la $out/*.ncu-rep $out/launches_${tag}.csv

#!/bin/bash
# round-2 session 13 (1 GPU): the driver's GPU tier + smoke on the final tree
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/r02_final_pytest_gpu.log 2>&1
tail -4 $out/r02_final_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/r02_final_smoke.log 2>&1; tail -1 $out/r02_final_smoke.log

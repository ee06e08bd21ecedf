#!/bin/bash
# round-2 session 14 (1 GPU, short): QR with L2-aware outer panels
out=gpurun_out
mkdir -p $out
timeout 60 python tools/qr_prof.py > $out/r02_qr_prof_l2aware.log 2>&1; cut -c1-220 $out/r02_qr_prof_l2aware.log
timeout 120 python -m pytest tests -m gpu -x -q -k "qr or svd_2048 or size_parity or canon" > $out/r02_qr_tests.log 2>&1; tail -3 $out/r02_qr_tests.log

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_train_exact.py 8192 16384 > gpurun_out/r_time.log 2>&1; echo "rc=$?" >> gpurun_out/r_time.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lstm_bwd_x3|lstm_fwd_x3' -s 4 -c 4 -o gpurun_out/r_prof_x3half -f python scripts/time_train_exact.py 8192 --noprof > gpurun_out/r_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/r_ncu.log

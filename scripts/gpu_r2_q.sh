#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_train_step.py 8192 16384 > gpurun_out/q_time.log 2>&1; echo "rc=$?" >> gpurun_out/q_time.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lstm_bwd_bf16|lstm2_fwd_train' -s 3 -c 3 -o gpurun_out/q_prof_half -f python scripts/time_train_step.py 8192 > gpurun_out/q_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/q_ncu.log

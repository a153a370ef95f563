#!/bin/bash
# final ncu captures of the training kernels (each only after its command has exited 0 without ncu)
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_train_step.py 8192 16384 > gpurun_out/u_time16.log 2>&1; echo "rc=$?" >> gpurun_out/u_time16.log
timeout 120 python scripts/time_train_exact.py 8192 16384 --noprof > gpurun_out/u_timex3.log 2>&1; echo "rc=$?" >> gpurun_out/u_timex3.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lstm_bwd_bf16|lstm2_fwd_train' -s 3 -c 6 -o gpurun_out/u_prof_tc -f python scripts/time_train_step.py 8192 16384 > gpurun_out/u_ncu16.log 2>&1; echo "rc=$?" >> gpurun_out/u_ncu16.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'lstm_bwd_x3|lstm_fwd_x3|lstm_wgrad_x3' -s 6 -c 12 -o gpurun_out/u_prof_x3 -f python scripts/time_train_exact.py 8192 16384 --noprof > gpurun_out/u_ncux3.log 2>&1; echo "rc=$?" >> gpurun_out/u_ncux3.log

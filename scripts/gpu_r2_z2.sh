#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/time_train_exact.py 8192 > gpurun_out/z2_x3.log 2>&1; echo "rc=$?" >> gpurun_out/z2_x3.log
timeout 120 python scripts/time_train_step.py 8192 > gpurun_out/z2_tc.log 2>&1; echo "rc=$?" >> gpurun_out/z2_tc.log

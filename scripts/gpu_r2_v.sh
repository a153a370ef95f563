#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/time_train_cpu.py 8192 > gpurun_out/v_cpu.log 2>&1; echo "rc=$?" >> gpurun_out/v_cpu.log

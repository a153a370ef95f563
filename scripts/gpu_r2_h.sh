#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -q --timeout 300 -x -k "wide_training" > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log
timeout 600 python scripts/time_stress_train.py 512 500 > gpurun_out/h_time_small.log 2>&1
timeout 900 python scripts/time_stress_train.py 2048 2500 > gpurun_out/h_time.log 2>&1

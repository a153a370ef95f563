#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q --timeout 600 -x > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
timeout 300 python scripts/time_train_half.py 64 2048 8192 9472 > gpurun_out/g_time.log 2>&1

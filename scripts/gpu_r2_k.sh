#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 300 python scripts/time_train_exact.py 16384 8192 > gpurun_out/k_time.log 2>&1
bash scripts/gpu_r2_j.sh

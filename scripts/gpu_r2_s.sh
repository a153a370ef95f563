#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 90 -x -k "exact_tc_training or gradients or five_class or several_tiles" > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 120 python scripts/time_train_exact.py 8192 16384 > gpurun_out/s_time.log 2>&1; echo "rc=$?" >> gpurun_out/s_time.log

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_bf16.py -m gpu -q --timeout 60 -x -k "train or grad or half or several" > gpurun_out/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p_pytest.log
timeout 120 python -m pytest tests/test_optim.py -m gpu -q --timeout 60 > gpurun_out/p_optim.log 2>&1; echo "rc=$?" >> gpurun_out/p_optim.log
timeout 120 python scripts/time_train_step.py 16384 8192 18944 > gpurun_out/p_time.log 2>&1; echo "rc=$?" >> gpurun_out/p_time.log

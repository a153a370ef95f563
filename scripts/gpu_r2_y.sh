#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/soak_train.py 1500 2 > gpurun_out/y_soak.log 2>&1; echo "rc=$?" >> gpurun_out/y_soak.log
timeout 200 python -m pytest tests/test_gpu_soak.py -m gpu -q > gpurun_out/y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/y_pytest.log

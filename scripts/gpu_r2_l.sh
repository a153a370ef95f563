#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_train_exact.py 32 8192 --noprof > gpurun_out/l_time_half.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_trainer.py -m gpu -q --timeout 60 -x > gpurun_out/l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/l_pytest.log
bash scripts/gpu_r2_j.sh

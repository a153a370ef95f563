#!/bin/bash
# validation of the x3 BPTT restructure + x3 half tiles + IIR half-block chains (every command under its own timeout)
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -x > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 120 python -m pytest tests/test_filters.py tests/test_gpu_bf16.py -m gpu -q --timeout 60 -x -k "filter or half or wide_training" > gpurun_out/k_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest2.log
timeout 120 python scripts/time_train_exact.py 16384 8192 --noprof > gpurun_out/k_time.log 2>&1
timeout 120 python scripts/time_iir_occ.py > gpurun_out/k_iir.log 2>&1

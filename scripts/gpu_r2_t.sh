#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -m gpu -q --timeout 90 -k "several_tiles" > gpurun_out/t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/t_pytest.log

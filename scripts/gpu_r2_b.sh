#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_x3_train.py 200 40 > gpurun_out/b_dbg1.log 2>&1; echo "rc=$?" >> gpurun_out/b_dbg1.log
timeout 300 python scripts/debug_x3_train.py 300 33 --drop > gpurun_out/b_dbg2.log 2>&1; echo "rc=$?" >> gpurun_out/b_dbg2.log
timeout 300 python scripts/time_filters.py > gpurun_out/b_filters.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "filter or five_class or stress or odd_sizes or fp16_tier" > gpurun_out/b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest.log

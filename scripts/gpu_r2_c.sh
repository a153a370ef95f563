#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_x3_train.py 300 33 > gpurun_out/c_dbg1.log 2>&1; echo "rc=$?" >> gpurun_out/c_dbg1.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-stress > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "rc=$?" >> gpurun_out/c_bench.err
ncu --set full --clock-control none -k regex:iir_chain -s 1 -c 1 -o gpurun_out/c_prof_iir -f python scripts/time_filters.py > gpurun_out/c_ncu_iir.log 2>&1

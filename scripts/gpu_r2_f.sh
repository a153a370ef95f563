#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase_filter.py tests/test_filters.py -m gpu -q --timeout 600 > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-stress --no-train > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?" >> gpurun_out/f_bench.err

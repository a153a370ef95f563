#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_optim.py tests/test_trainer.py tests/test_gpu_bf16.py tests/test_gpu_parity.py -m gpu -q --timeout 90 -x > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-stress --no-cpu > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "rc=$?" >> gpurun_out/o_bench.err

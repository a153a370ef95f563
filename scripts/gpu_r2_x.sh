#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 280 > gpurun_out/x_pytest_dp.log 2>&1; echo "rc=$?" >> gpurun_out/x_pytest_dp.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-stress --no-cpu > gpurun_out/n2b_bench.json 2> gpurun_out/n2b_bench.err; echo "rc=$?" >> gpurun_out/n2b_bench.err

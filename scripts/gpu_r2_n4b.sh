#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 --no-stress --no-cpu > gpurun_out/n4b_bench.json 2> gpurun_out/n4b_bench.err; echo "rc=$?" >> gpurun_out/n4b_bench.err

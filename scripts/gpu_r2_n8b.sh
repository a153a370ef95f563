#!/bin/bash
# 8-GPU bench run only (the reference arm and the DP pytest ran earlier in the round)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-stress --no-cpu > gpurun_out/n8b_bench.json 2> gpurun_out/n8b_bench.err; echo "rc=$?" >> gpurun_out/n8b_bench.err

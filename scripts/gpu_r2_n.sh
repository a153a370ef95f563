#!/bin/bash
# multi-GPU bench run: bash scripts/gpu_r2_n.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-stress > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "rc=$?" >> gpurun_out/n${N}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/n${N}_ref.json 2> gpurun_out/n${N}_ref.err; echo "rc=$?" >> gpurun_out/n${N}_ref.err
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 600 > gpurun_out/n2_pytest_dp.log 2>&1; echo "rc=$?" >> gpurun_out/n2_pytest_dp.log; fi

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/diag_train_n.py 8192 > gpurun_out/w_diag.log 2>&1; echo "rc=$?" >> gpurun_out/w_diag.log

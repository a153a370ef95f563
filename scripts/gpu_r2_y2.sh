#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/soak_train.py 5000 11 > gpurun_out/y2_soak.log 2>&1; echo "rc=$?" >> gpurun_out/y2_soak.log

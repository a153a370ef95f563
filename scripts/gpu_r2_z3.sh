#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu > gpurun_out/z3_bench.json 2> gpurun_out/z3_bench.err; echo "rc=$?" >> gpurun_out/z3_bench.err

#!/bin/bash
# final single-GPU validation: what the driver runs at round end (every command under its own timeout)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; echo "rc=$?" >> gpurun_out/z_ref.err
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "rc=$?" >> gpurun_out/z_bench.err

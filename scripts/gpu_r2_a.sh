#!/bin/bash
# GPU session A of round 2: parity tests, smoke, bench, sanitizer logs, ncu of the training kernels and the headline kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/a_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
python scripts/time_filters.py > gpurun_out/a_filters.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/a_racecheck.log
# ncu: final training kernels (16-bit tier) + headline x3 kernel at the benchmark size
python scripts/time_train_step.py 16384 > gpurun_out/a_train_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'lstm2_fwd_train_v2|lstm_bwd_bf16' -s 6 -c 3 -o gpurun_out/a_prof_train -f python scripts/time_train_step.py 16384 > gpurun_out/a_ncu_train.log 2>&1
python scripts/prof_x3.py 40960 > gpurun_out/a_x3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:decoder_infer_x3 -s 1 -c 1 -o gpurun_out/a_prof_x3 -f python scripts/prof_x3.py 40960 > gpurun_out/a_ncu_x3.log 2>&1
ncu --set full --clock-control none -k regex:iir_chain -s 1 -c 1 -o gpurun_out/a_prof_iir -f python scripts/time_filters.py > gpurun_out/a_ncu_iir.log 2>&1
ls -la gpurun_out | tail -20

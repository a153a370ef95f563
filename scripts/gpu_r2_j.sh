#!/bin/bash
# evidence session: ncu --set full of the round-2 kernels + launch list of the bench command
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_train_exact.py 16384 --noprof > gpurun_out/j_exact_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lstm_fwd_x3|lstm_bwd_x3|lstm_wgrad_x3' -s 6 -c 6 -o gpurun_out/j_prof_x3train -f python scripts/time_train_exact.py 16384 --noprof > gpurun_out/j_ncu_x3train.log 2>&1
timeout 120 python scripts/prof_phase_iir.py > gpurun_out/j_phase_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none -k regex:'iir_chain|phase_coupling' -s 2 -c 2 -o gpurun_out/j_prof_phase_iir -f python scripts/prof_phase_iir.py > gpurun_out/j_ncu_phase.log 2>&1
timeout 120 python scripts/time_train_half.py 8192 > gpurun_out/j_half_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none -k regex:'lstm2_fwd_train_v2|lstm_bwd_bf16' -s 9 -c 3 -o gpurun_out/j_prof_half -f python scripts/time_train_half.py 8192 > gpurun_out/j_ncu_half.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:lstm_wide -s 0 -c 4 -o gpurun_out/j_prof_wide -f python scripts/time_stress_train.py 2048 60 > gpurun_out/j_ncu_wide.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-stress --no-cpu > gpurun_out/j_bench_plain.json 2> gpurun_out/j_bench_plain.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 2 --warmup 3 --no-stress --no-cpu > gpurun_out/j_ncu_bench.log 2>&1
ls -la gpurun_out | grep " j_"

#!/bin/bash
# launch list + full ncu captures of the step kernels and the persistent trust-region search (1 GPU, C2)
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches12_c2.csv python bench.py --steps 130 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu12_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell|k_primal_step|k_step_decide' -s 60 -c 8 -o gpurun_out/prof12_step python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu12_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tr_search|k_tr_prepare' -s 4 -c 2 -o gpurun_out/prof12_tr python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu12_tr.log 2>&1
ls -la gpurun_out/*.ncu-rep

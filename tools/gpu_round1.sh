#!/bin/bash
# One GPU-box session: parity tests, bench, launch list, full ncu capture, layout probes.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --scale 0.05 --steps 200 --warmup 16 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -2 gpurun_out/bench_small.err
timeout 900 python bench.py --steps 1000 --warmup 64 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -5 gpurun_out/bench_c2.err
cat gpurun_out/bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell -s 40 -c 4 -o gpurun_out/prof_sell python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_sell.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_primal_step -s 5 -c 1 -o gpurun_out/prof_primal python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_primal.log 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/spmv_probe.cu -o /tmp/spmv_probe && timeout 300 /tmp/spmv_probe > gpurun_out/spmv_probe.txt 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/gather_probe.cu -o /tmp/gather_probe && timeout 300 /tmp/gather_probe > gpurun_out/gather_probe.txt 2>&1
tail -40 gpurun_out/spmv_probe.txt; tail -30 gpurun_out/gather_probe.txt

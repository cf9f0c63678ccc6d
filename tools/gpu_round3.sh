#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_device_build.py -m gpu -x -q > gpurun_out/pytest_build3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_build3.log; tail -30 gpurun_out/pytest_build3.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log; tail -15 gpurun_out/pytest_gpu3.log
PDLP_B200_TRACE=1 timeout 900 python bench.py --steps 1000 --warmup 64 > gpurun_out/bench3_c2_n1.json 2> gpurun_out/bench3_c2_n1.err; tail -4 gpurun_out/bench3_c2_n1.err; cat gpurun_out/bench3_c2_n1.json | cut -c1-3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches3_c2.csv python bench.py --steps 130 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_bench3.log 2>&1

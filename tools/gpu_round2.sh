#!/bin/bash
# 2-GPU session: all parity tests (incl. row-sharded), bench at N=1 and N=2, launch list.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
PDLP_B200_TRACE=1 timeout 900 python bench.py --steps 1000 --warmup 64 > gpurun_out/bench2_c2_n1.json 2> gpurun_out/bench2_c2_n1.err; tail -12 gpurun_out/bench2_c2_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 64 > gpurun_out/bench2_c2_n2.json 2> gpurun_out/bench2_c2_n2.err; tail -5 gpurun_out/bench2_c2_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --scale 0.25 --steps 500 --warmup 64 --no-cpu > gpurun_out/bench2_c4q_n2.json 2> gpurun_out/bench2_c4q_n2.err; tail -5 gpurun_out/bench2_c4q_n2.err
timeout 900 python bench.py --config c4 --scale 0.25 --steps 500 --warmup 64 --no-cpu > gpurun_out/bench2_c4q_n1.json 2> gpurun_out/bench2_c4q_n1.err; tail -3 gpurun_out/bench2_c4q_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches2_c2.csv python bench.py --steps 130 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_bench2.log 2>&1
cat gpurun_out/bench2_c2_n1.json gpurun_out/bench2_c2_n2.json gpurun_out/bench2_c4q_n2.json gpurun_out/bench2_c4q_n1.json | cut -c1-1500

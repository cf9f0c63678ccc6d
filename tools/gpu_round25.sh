#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q -k "peer-d" > gpurun_out/pytest_gpu25.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu25.log; tail -4 gpurun_out/pytest_gpu25.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --steps 300"
timeout 600 $T > gpurun_out/bench25_n2.json 2> gpurun_out/bench25_n2.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench25_n2.json').read().splitlines() if l.startswith('{')][-1])
print(d['value'], d['config']['parallelism'], d['e2e']['value'], [k['kernel'][:30] for k in d['kernels']])"

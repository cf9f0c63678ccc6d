#!/bin/bash
# N-GPU record run: parity of the sharded paths + device build (pooled temporaries), bench at N
set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m pytest tests/test_device_build.py tests/test_distributed.py tests/test_boundary.py -m gpu -x -q > gpurun_out/pytest_gpu19.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu19.log; tail -4 gpurun_out/pytest_gpu19.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 64 --no-cpu"
PDLP_B200_TRACE=1 timeout 900 $T > gpurun_out/bench19_n$N.json 2> gpurun_out/bench19_n$N.err; grep "trace\] \(step\|entry\)" gpurun_out/bench19_n$N.err | head -3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench19_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

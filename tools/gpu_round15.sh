#!/bin/bash
# all-gather exchange (peer-d): parity of the three exchanges, bench A/B peer-d vs peer-s
set -x
mkdir -p gpurun_out
N=${1:-2}
[ "$2" = "notest" ] || timeout 1200 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu15.log; tail -25 gpurun_out/pytest_gpu15.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 600 --warmup 64 --no-cpu"
for ex in peer-d peer-s; do
PDLP_B200_EXCHANGE=$ex PDLP_B200_TRACE=1 timeout 900 $T > gpurun_out/bench15_${ex}_n$N.json 2> gpurun_out/bench15_${ex}_n$N.err; grep "trace\]" gpurun_out/bench15_${ex}_n$N.err | grep -v arena | head -3
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench15_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

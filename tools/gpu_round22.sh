#!/bin/bash
# final N-GPU record (default exchange selection)
set -x
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
PDLP_B200_TRACE=1 timeout 900 $T > gpurun_out/bench22_n$N.json 2> gpurun_out/bench22_n$N.err; grep "trace\] step" gpurun_out/bench22_n$N.err | head -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench22_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

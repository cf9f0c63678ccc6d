#!/bin/bash
# persistent trust-region search: parity (1 GPU + 2 GPU) and bench A/B vs the legacy multi-launch search
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.log; tail -5 gpurun_out/pytest_gpu11.log
PDLP_B200_TRACE=1 timeout 600 python bench.py --steps 600 --warmup 64 --no-cpu > gpurun_out/bench11_n1.json 2> gpurun_out/bench11_n1.err; grep trace gpurun_out/bench11_n1.err | head -3
PDLP_B200_TR_LEGACY=1 PDLP_B200_TRACE=1 timeout 600 python bench.py --steps 600 --warmup 64 --no-cpu --no-e2e > gpurun_out/bench11_n1_legacy.json 2> gpurun_out/bench11_n1_legacy.err; grep trace gpurun_out/bench11_n1_legacy.err | head -3
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 600 --warmup 64 --no-cpu"
PDLP_B200_TRACE=1 timeout 900 $T > gpurun_out/bench11_peer_n$N.json 2> gpurun_out/bench11_peer_n$N.err; grep "trace\]" gpurun_out/bench11_peer_n$N.err | head -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench11_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

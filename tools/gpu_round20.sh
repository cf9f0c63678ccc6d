#!/bin/bash
# C4 (200 M nnz, tall) on N GPUs
set -x
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
PDLP_B200_TRACE=1 timeout 1500 python bench.py --config c4 --steps 200 --warmup 64 --no-cpu --no-e2e > gpurun_out/bench20_c4_n1.json 2> gpurun_out/bench20_c4_n1.err; grep -v "^+" gpurun_out/bench20_c4_n1.err | tail -5
else
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config c4 --gpus $N --steps 200 --warmup 64 --no-cpu --no-e2e"
PDLP_B200_TRACE=1 timeout 1500 $T > gpurun_out/bench20_c4_n$N.json 2> gpurun_out/bench20_c4_n$N.err; grep "trace\] step\|bench\]" gpurun_out/bench20_c4_n$N.err | head -6
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench20_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'frac', d['iteration_roofline']['frac_of_peak'])
    except Exception as e:
        print(f,'ERR',e)
PY

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernel_goldens.py tests/test_solver_goldens.py -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu13.log; tail -3 gpurun_out/pytest_gpu13.log
PDLP_B200_TRACE=1 timeout 600 python bench.py --steps 600 --warmup 64 --no-cpu > gpurun_out/bench13_n1.json 2> gpurun_out/bench13_n1.err; grep trace gpurun_out/bench13_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench13_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'), (d.get('e2e') or {}).get('wall_s'))
    except Exception as e:
        print(f,'ERR',e)
PY

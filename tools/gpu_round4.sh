#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 600 --warmup 64 --no-e2e --no-cpu"
for v in 6 0 8 4; do
  PDLP_B200_SELL_BLOCKS_PER_SM=$v timeout 600 $B > gpurun_out/ab4_persist$v.json 2> gpurun_out/ab4_persist$v.err
done
PDLP_B200_HOST_BUILD=1 timeout 600 $B > gpurun_out/ab4_hostbuild.json 2> gpurun_out/ab4_hostbuild.err
timeout 900 python bench.py --config c3 --steps 600 --warmup 64 > gpurun_out/bench4_c3.json 2> gpurun_out/bench4_c3.err; tail -3 gpurun_out/bench4_c3.err
timeout 900 python bench.py --config c5 --steps 600 --warmup 64 > gpurun_out/bench4_c5.json 2> gpurun_out/bench4_c5.err; tail -3 gpurun_out/bench4_c5.err
timeout 1200 python bench.py --config c4 --steps 300 --warmup 64 --no-cpu > gpurun_out/bench4_c4.json 2> gpurun_out/bench4_c4.err; tail -3 gpurun_out/bench4_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab4_*.json')+glob.glob('gpurun_out/bench4_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'step-only frac %.3f'%d['iteration_roofline']['step_loop_only_frac'], ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY

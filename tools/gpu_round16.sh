#!/bin/bash
# 1-GPU validation: full GPU suite, C2 bench (with CPU baseline + reference arm), C3 / C5 benches
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu16.log; tail -4 gpurun_out/pytest_gpu16.log
PDLP_B200_TRACE=1 timeout 900 python bench.py --steps 1000 --warmup 64 > gpurun_out/bench16_c2.json 2> gpurun_out/bench16_c2.err; grep trace gpurun_out/bench16_c2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench16_ref.json 2> gpurun_out/bench16_ref.err; tail -2 gpurun_out/bench16_ref.err; cat gpurun_out/bench16_ref.json | cut -c1-600
timeout 900 python bench.py --config c3 --steps 600 --warmup 64 --no-cpu > gpurun_out/bench16_c3.json 2> gpurun_out/bench16_c3.err
timeout 900 python bench.py --config c5 --steps 600 --warmup 64 --no-cpu > gpurun_out/bench16_c5.json 2> gpurun_out/bench16_c5.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke16.log 2>&1; tail -2 gpurun_out/smoke16.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench16_c*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'frac', d['roofline']['frac'], d['iteration_roofline']['frac_of_peak'], 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY

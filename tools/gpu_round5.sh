#!/bin/bash
# Session 6: full GPU parity suite on the persistent SELL kernels, A/B of the grid cap,
# C3/C5 benches, full ncu capture of the two fused SELL kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi5.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log; tail -5 gpurun_out/pytest_gpu5.log
B="python bench.py --steps 600 --warmup 64 --no-e2e --no-cpu"
for v in 6 0; do
  PDLP_B200_SELL_BLOCKS_PER_SM=$v timeout 600 $B > gpurun_out/ab5_persist$v.json 2> gpurun_out/ab5_persist$v.err
done
timeout 900 python bench.py --steps 1000 --warmup 64 > gpurun_out/bench5_c2.json 2> gpurun_out/bench5_c2.err; tail -3 gpurun_out/bench5_c2.err
timeout 900 python bench.py --config c3 --steps 600 --warmup 64 --no-cpu > gpurun_out/bench5_c3.json 2> gpurun_out/bench5_c3.err; tail -3 gpurun_out/bench5_c3.err
timeout 900 python bench.py --config c5 --steps 600 --warmup 64 --no-cpu > gpurun_out/bench5_c5.json 2> gpurun_out/bench5_c5.err; tail -3 gpurun_out/bench5_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell -s 40 -c 4 -o gpurun_out/prof5_sell python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu5_full_sell.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab5_*.json')+glob.glob('gpurun_out/bench5_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'step-only frac %.3f'%d['iteration_roofline']['step_loop_only_frac'], ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY

#!/bin/bash
set -x
mkdir -p gpurun_out
PDLP_B200_TRACE=1 timeout 900 python bench.py --steps 1000 --warmup 64 --no-cpu > gpurun_out/bench18_a.json 2> gpurun_out/bench18_a.err; grep "trace\] \(SELL\|entry\)" gpurun_out/bench18_a.err
timeout 900 python bench.py > gpurun_out/bench18_b.json 2> gpurun_out/bench18_b.err; tail -3 gpurun_out/bench18_b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench18_*.json')):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    print(f, d['value'], 'e2e', d['e2e']['value'], d['e2e']['iterations'], d['e2e']['wall_s'], (d.get('cpu_baseline') or {}).get('value'))
PY

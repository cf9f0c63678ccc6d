#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2 3; do
PDLP_B200_TRACE=1 timeout 600 python bench.py --steps 64 --warmup 3 --no-cpu > gpurun_out/bench17_e2e$i.json 2> gpurun_out/bench17_e2e$i.err; grep "trace\] \(SELL\|entry\)" gpurun_out/bench17_e2e$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench17_*.json')):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    print(f, 'e2e', d['e2e']['value'], d['e2e']['iterations'], d['e2e']['wall_s'])
PY

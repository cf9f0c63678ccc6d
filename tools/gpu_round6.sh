#!/bin/bash
# Session 6b: A/B of SELL kernel variants (block size x software pipelining), parity of each variant.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 600 --warmup 64 --no-e2e --no-cpu"
for v in 0 1 2; do for t in 128 256 512; do
  PDLP_B200_SELL_VARIANT=$v PDLP_B200_SELL_THREADS=$t timeout 600 $B > gpurun_out/ab6_v${v}_t${t}.json 2> gpurun_out/ab6_v${v}_t${t}.err
done; done
for v in 1 2; do
PDLP_B200_SELL_VARIANT=$v PDLP_B200_SELL_THREADS=128 timeout 900 python -m pytest tests/test_kernel_goldens.py tests/test_synthetic_configs.py -m gpu -x -q > gpurun_out/pytest_gpu6_v$v.log 2>&1; tail -3 gpurun_out/pytest_gpu6_v$v.log
done
PDLP_B200_TRACE=1 timeout 600 python bench.py --steps 200 --warmup 64 --no-cpu > gpurun_out/bench6_trace.json 2> gpurun_out/bench6_trace.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab6_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'step-only frac %.3f'%d['iteration_roofline']['step_loop_only_frac'], ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']))
    except Exception as e:
        print(f,'ERR',e)
PY

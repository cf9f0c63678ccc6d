#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu26.log; tail -5 gpurun_out/pytest_gpu26.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu > gpurun_out/bench26.json 2> gpurun_out/bench26.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench26.json').read().splitlines() if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])"

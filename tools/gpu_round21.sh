#!/bin/bash
# final 1-GPU record: full GPU suite, smoke, default bench (+ reference arm), launch list, full ncu captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu21.log; tail -4 gpurun_out/pytest_gpu21.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke21.log 2>&1; tail -2 gpurun_out/smoke21.log
timeout 900 python bench.py > gpurun_out/bench21_c2.json 2> gpurun_out/bench21_c2.err; tail -2 gpurun_out/bench21_c2.err
timeout 600 python bench.py --impl reference > gpurun_out/bench21_ref.json 2> gpurun_out/bench21_ref.err; cut -c1-300 gpurun_out/bench21_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches21_c2.csv python bench.py --steps 130 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu21_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell|k_primal_step|k_step_decide' -s 60 -c 4 -o gpurun_out/prof21_step python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu21_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tr_search' -s 2 -c 1 -o gpurun_out/prof21_tr python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu21_tr.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench21_c*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'frac', d['roofline']['frac'], d['roofline']['traffic'], d['iteration_roofline']['frac_of_peak'], 'e2e', (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY

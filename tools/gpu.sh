#!/bin/bash
# One entry point for the GPU-box jobs of a round (run under gpurun from the repo root):
#   tools/gpu.sh <tag> <stage> [<stage> ...]
# Everything a stage writes goes to gpurun_out/<tag>_*; summaries worth keeping are copied to
# profiles/ by hand afterwards. Stages:
#   sanitize    compute-sanitizer memcheck / racecheck / initcheck / synccheck (tools/sanitize.sh)
#   ncu-full    ncu --set full on LIVE step-loop / trust-region launches of C2 (warm-up skipped)
#   launches    ncu launch list (gpu__time_duration) of a short C2 bench
#   bench       default bench.py (1000 steps) and the driver's window (--steps 20 --warmup 5)
#   bench-cfgs  C3 / C5 lines
#   pytest      the whole -m gpu suite
#   pytest-big  only the full-size parity tests
#   ab          A/B of the kernel variants behind env knobs (PDLP_B200_*), see body
set -u
cd "$(dirname "$0")/.."
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    k=d.get('kernels') or []
    print(sys.argv[1], 'value %.1f'%d['value'], 'loop_ms %.2f dev_ms/it %.4f'%(d.get('device_step_loop_ms',0), d['ms_per_step']),
          'kern_us', ' '.join('%.1f'%(1000*(x['avg_ms'] or 0)) for x in k), 'frac', '%.3f'%(d['roofline']['frac'] or 0),
          'iter_frac %.3f/%.3f'%(d['iteration_roofline']['frac_of_peak'], d['iteration_roofline']['step_loop_only_frac'] or 0),
          'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for stage in "$@"; do
  echo "===== stage $stage"
  case $stage in
    sanitize)
      timeout 1500 tools/sanitize.sh > $OUT/${TAG}_sanitize.log 2>&1; echo "sanitize rc=$?"; grep -E "^===|ERROR SUMMARY|RACECHECK SUMMARY" $OUT/${TAG}_sanitize.log ;;
    ncu-full)
      # warm-up launches (problem build, rescaling, first iterations) are skipped so only live step launches are captured;
      # the reports are summarised ON THE BOX (gpurun_out is capped at 64 MiB) and only the CSVs travel back
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell|k_primal_step' -s 60 -c 9 \
        -o /tmp/${TAG}_step -f python bench.py --steps 40 --warmup 8 --no-e2e --no-cpu > $OUT/${TAG}_ncu_step.log 2>&1; echo "ncu step rc=$?"
      python tools/ncu_summary.py /tmp/${TAG}_step.ncu-rep $OUT/${TAG}_ncu_full_step_c2.csv --traffic-out $OUT/${TAG}_traffic.json c2
      ncu -i /tmp/${TAG}_step.ncu-rep --page source --csv --kernel-name regex:DualEpi 2>/dev/null | head -700 > $OUT/${TAG}_ncu_source_dualepi.csv
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tr_' -s 4 -c 2 \
        -o /tmp/${TAG}_tr -f python bench.py --steps 40 --warmup 8 --no-e2e --no-cpu > $OUT/${TAG}_ncu_tr.log 2>&1; echo "ncu tr rc=$?"
      python tools/ncu_summary.py /tmp/${TAG}_tr.ncu-rep $OUT/${TAG}_ncu_full_tr_c2.csv
      # the TMA-staged variant of the same step kernels, for the counter comparison (NCU_TMA=0 skips it)
      [ "${NCU_TMA:-1}" = "1" ] && PDLP_B200_SELL_VARIANT=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell_tma' -s 60 -c 6 \
        -o /tmp/${TAG}_step_tma -f python bench.py --steps 40 --warmup 8 --no-e2e --no-cpu > $OUT/${TAG}_ncu_step_tma.log 2>&1; echo "ncu tma rc=$?"
      [ "${NCU_TMA:-1}" = "1" ] && python tools/ncu_summary.py /tmp/${TAG}_step_tma.ncu-rep $OUT/${TAG}_ncu_full_step_tma_c2.csv
      ls -la $OUT/${TAG}_ncu_full_*.csv ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/${TAG}_launches_raw.csv \
        python bench.py --steps 130 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
      python tools/compact_launches.py $OUT/${TAG}_launches_raw.csv $OUT/${TAG}_launches.csv "bench.py --steps 130 --warmup 3 (C2, 1 B200), ncu gpu__time_duration.sum --clock-control none"; head -30 $OUT/${TAG}_launches.csv ;;
    launches-driver)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/${TAG}_launches_driver_raw.csv \
        python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > $OUT/${TAG}_launches_driver.log 2>&1; echo "launch list rc=$?"
      python tools/compact_launches.py $OUT/${TAG}_launches_driver_raw.csv $OUT/${TAG}_launches_driver.csv "bench.py --steps 20 --warmup 5 (C2, 1 B200), ncu gpu__time_duration.sum --clock-control none"; head -40 $OUT/${TAG}_launches_driver.csv ;;
    bench)
      PDLP_B200_TRACE=1 timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; line $OUT/${TAG}_bench.json
      grep "trace\]" $OUT/${TAG}_bench.err | head -8
      timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/${TAG}_bench_driver.json 2> $OUT/${TAG}_bench_driver.err; line $OUT/${TAG}_bench_driver.json ;;
    bench-cfgs)
      for c in c3 c5; do timeout 900 python bench.py --config $c --no-cpu > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; line $OUT/${TAG}_bench_$c.json; done ;;
    pytest)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log ;;
    pytest-big)
      timeout 900 python -m pytest tests -m gpu -x -q -k "full_size" > $OUT/${TAG}_pytest_big.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_big.log ;;
    ab)
      for v in ${AB_VARIANTS:-"default"}; do
        env $(echo $v | tr ',' ' ') timeout 300 python bench.py --no-e2e --no-cpu > $OUT/${TAG}_ab_$(echo $v | tr -c 'A-Za-z0-9_\n' '_').json 2>/dev/null
        echo "-- $v"; line $OUT/${TAG}_ab_$(echo $v | tr -c 'A-Za-z0-9_\n' '_').json
      done ;;
    pytest-dist)   # multi-GPU parity tests (run under gpurun --gpus N); the log is kept under profiles/ by hand
      timeout 1700 python -m pytest tests/test_distributed.py -m gpu -q -rs > $OUT/${TAG}_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest_dist.log ;;
    bench-n)       # bench.py on NGPU GPUs (driver window) with the per-phase trace; C4 sub-record at C4_SCALE
      T="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2}"
      PDLP_B200_TRACE=1 timeout 1500 $T --steps 20 --warmup 5 --c4-scale ${C4_SCALE:-1.0} ${BENCH_EXTRA:-} > $OUT/${TAG}_bench_n${NGPU:-2}.json 2> $OUT/${TAG}_bench_n${NGPU:-2}.err
      line $OUT/${TAG}_bench_n${NGPU:-2}.json
      grep -E "trace\] (step|rank 0|SELL|entry)|\[bench\]" $OUT/${TAG}_bench_n${NGPU:-2}.err | head -24
      python - $OUT/${TAG}_bench_n${NGPU:-2}.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print('c4', json.dumps(d.get('c4')))
print('e2e_cold', json.dumps(d.get('e2e_cold')))
PY
      ;;
    *) echo "unknown stage $stage" ;;
  esac
done

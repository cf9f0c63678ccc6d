set -u
cd /root/repo
for t in 0 1; do
  PDLP_B200_TEAM_SLOTS=$t timeout 600 python bench.py --config c3 --no-cpu --no-e2e > gpurun_out/r02u_bench_c3_team$t.json 2>/dev/null
  PDLP_B200_TEAM_SLOTS=$t timeout 600 ncu --set full --clock-control none -k regex:k_sell -s 40 -c 4 -o /tmp/c3_team$t -f python bench.py --config c3 --steps 30 --warmup 8 --no-e2e --no-cpu > gpurun_out/r02u_ncu_c3_team$t.log 2>&1
  python tools/ncu_summary.py /tmp/c3_team$t.ncu-rep gpurun_out/r02u_ncu_full_step_c3_team$t.csv
done
python - <<'PY'
import json,csv
for t in (0,1):
    d=json.loads([l for l in open('gpurun_out/r02u_bench_c3_team%d.json'%t).read().splitlines() if l.startswith('{')][-1])
    print('team',t,'value %.1f'%d['value'],'kern_us',' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']))
    rows=[r for r in csv.reader(l for l in open('gpurun_out/r02u_ncu_full_step_c3_team%d.csv'%t) if not l.startswith('#'))]
    h=rows[0]; idx={n.split(' [')[0]:i for i,n in enumerate(h)}
    for r in rows[1:]:
        print('   ',r[0][:44],'us',r[idx['gpu__time_duration.sum']][:7],'l1hit',r[idx['l1tex__t_sector_hit_rate.pct']][:6],'sectors',r[idx['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']][:10],'req',r[idx['l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']][:9],'xbar%',r[idx['l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed']][:5])
PY
timeout 900 python -m pytest tests/test_device_build.py tests/test_kernel_goldens.py tests/test_synthetic_configs.py -m gpu -x -q 2>&1 | tail -3

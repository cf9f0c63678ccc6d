#!/bin/bash
# A/B: programmatic dependent launch, slot groups per block; parity of the defaults; 2-GPU check
set -x
mkdir -p gpurun_out
B="python bench.py --steps 600 --warmup 64 --no-e2e --no-cpu"
PDLP_B200_PDL=0 PDLP_B200_SELL_CHUNKS=1 timeout 600 $B > gpurun_out/ab14_pdl0_c1.json 2> gpurun_out/ab14_pdl0_c1.err
PDLP_B200_PDL=1 PDLP_B200_SELL_CHUNKS=1 timeout 600 $B > gpurun_out/ab14_pdl1_c1.json 2> gpurun_out/ab14_pdl1_c1.err
PDLP_B200_PDL=1 PDLP_B200_SELL_CHUNKS=2 timeout 600 $B > gpurun_out/ab14_pdl1_c2.json 2> gpurun_out/ab14_pdl1_c2.err
PDLP_B200_PDL=1 PDLP_B200_SELL_CHUNKS=4 timeout 600 $B > gpurun_out/ab14_pdl1_c4.json 2> gpurun_out/ab14_pdl1_c4.err
PDLP_B200_PDL=1 PDLP_B200_SELL_CHUNKS=8 timeout 600 $B > gpurun_out/ab14_pdl1_c8.json 2> gpurun_out/ab14_pdl1_c8.err
PDLP_B200_PDL=0 PDLP_B200_SELL_CHUNKS=4 timeout 600 $B > gpurun_out/ab14_pdl0_c4.json 2> gpurun_out/ab14_pdl0_c4.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu14.log; tail -3 gpurun_out/pytest_gpu14.log
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 600 --warmup 64 --no-cpu --no-e2e"
PDLP_B200_TRACE=1 timeout 900 $T > gpurun_out/ab14_peer_n$N.json 2> gpurun_out/ab14_peer_n$N.err; grep "trace\]" gpurun_out/ab14_peer_n$N.err | head -4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab14_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']))
    except Exception as e:
        print(f,'ERR',e)
PY

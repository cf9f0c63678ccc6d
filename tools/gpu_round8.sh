#!/bin/bash
# N-GPU session: parity (both exchanges) + bench of the peer exchange; NCCL exchange A/B when asked.
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_gpu8_peer.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu8_peer.log; tail -5 gpurun_out/pytest_gpu8_peer.log
PDLP_B200_EXCHANGE=nccl timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_gpu8_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu8_nccl.log; tail -5 gpurun_out/pytest_gpu8_nccl.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 600 --warmup 64 --no-cpu"
timeout 900 $T > gpurun_out/bench8_peer_n$N.json 2> gpurun_out/bench8_peer_n$N.err; tail -5 gpurun_out/bench8_peer_n$N.err
if [ "$2" = "ab" ]; then
PDLP_B200_EXCHANGE=nccl timeout 900 $T > gpurun_out/bench8_nccl_n$N.json 2> gpurun_out/bench8_nccl_n$N.err; tail -5 gpurun_out/bench8_nccl_n$N.err
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench8_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], 'loop ms %.1f wall %.1f'%(d['device_step_loop_ms'], d['wall_ms_timed']), ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

#!/bin/bash
# 2-GPU session: peer-memory exchange vs NCCL all-reduce exchange (parity test + bench A/B).
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo7.txt 2>&1
timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/pytest_gpu7_peer.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu7_peer.log; tail -15 gpurun_out/pytest_gpu7_peer.log
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 600 --warmup 64 --no-cpu"
timeout 900 $T > gpurun_out/bench7_peer_n$N.json 2> gpurun_out/bench7_peer_n$N.err; tail -5 gpurun_out/bench7_peer_n$N.err
PDLP_B200_EXCHANGE=nccl timeout 900 $T > gpurun_out/bench7_nccl_n$N.json 2> gpurun_out/bench7_nccl_n$N.err; tail -5 gpurun_out/bench7_nccl_n$N.err
timeout 600 python bench.py --steps 600 --warmup 64 --no-cpu > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench7_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f'%d['value'], ' '.join('%.1f'%(1000*(k['avg_ms'] or 0)) for k in d['kernels']), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('iterations'))
    except Exception as e:
        print(f,'ERR',e)
PY

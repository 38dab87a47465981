#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|rror" | head
for nb in 512 1024 2048; do
  GPB200_DIST_NB=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $NG --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('dist_nb=$nb gpus', j['n_gpus'], 'ms', round(j['ms_per_step'],1), 'GF', round(j['value']), j['config']['phases_ms'], 'roof', j['roofline']['frac'] if j['roofline'] else None)
"
done

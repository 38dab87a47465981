#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|rror" | head
for ng in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $ng --steps 3 --warmup 3 2>gpurun_out/bench_mg_$ng.err | tee gpurun_out/bench_mg_$ng.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('gpus', j['n_gpus'], 'ms', round(j['ms_per_step'],1), 'GF', round(j['value']), j['config']['phases_ms'], 'roof', j['roofline']['frac'] if j['roofline'] else None, 'e2e', round(j['e2e']['ms_per_step'],1))
"
done
GPB200_DIST_NB=512 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('nb512 gpus', j['n_gpus'], 'ms', round(j['ms_per_step'],1), j['config']['phases_ms'])
"

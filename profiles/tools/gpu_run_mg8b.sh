#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|rror" | head
run() { # ngpu nb p2p
  GPB200_DIST_NB=$2 GPB200_P2P=$3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $1 --steps 3 --warmup 3 2>gpurun_out/bench_mg.err | tee gpurun_out/bench_mg_$1_nb$2_p2p$3.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('gpus', j['n_gpus'], 'nb $2 p2p $3 ms', round(j['ms_per_step'],1), 'GF', round(j['value']), j['config']['phases_ms'], 'e2e', round(j['e2e']['ms_per_step'],1))
"
}
run 8 1024 1
run 8 512 1
run 8 256 1
run 8 512 0
run 4 1024 1

#!/bin/bash
mkdir -p gpurun_out
run() { # ngpu p2p(env or empty)
  if [ -n "$2" ]; then export GPB200_P2P=$2; else unset GPB200_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $1 --steps 3 --warmup 3 2>gpurun_out/bench_mg.err | tee gpurun_out/bench_mg_$1_p2p${2:-auto}.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('gpus', j['n_gpus'], 'p2p ${2:-auto} ms', round(j['ms_per_step'],1), 'GF', round(j['value']), j['config']['phases_ms'], 'e2e', round(j['e2e']['ms_per_step'],1))
"
}
run 4 0
run 4 1
run 2

#!/bin/bash
# multi-GPU bring-up on N GPUs (argument), default 2
NG=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 tests/mgpu_worker.py > gpurun_out/mgpu_worker_$NG.log 2>&1
echo "worker exit $?" >> gpurun_out/mgpu_worker_$NG.log
tail -n 12 gpurun_out/mgpu_worker_$NG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_mg_$NG.json 2> gpurun_out/bench_mg_$NG.err
echo "bench exit $?"; tail -n 5 gpurun_out/bench_mg_$NG.err; cat gpurun_out/bench_mg_$NG.json

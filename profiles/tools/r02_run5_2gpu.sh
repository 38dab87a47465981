#!/bin/bash
# round 2, call 5 (2 GPUs): NCCL paths -- replicated (fused P2P push / NCCL broadcast) and row-sharded (plain / look-ahead)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_2gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -s > gpurun_out/r02_pytest_multi_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_multi_2.log
grep -E "MGPU|passed|failed|skipped" gpurun_out/r02_pytest_multi_2.log | tail -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_repl.json 2> gpurun_out/r02_bench_2gpu_repl.err
GPB200_P2P=0 GPB200_SHARD=1 timeout 600 $TR --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_shard_la.json 2> gpurun_out/r02_bench_2gpu_shard_la.err
GPB200_P2P=0 GPB200_SHARD=1 GPB200_SHARD_LA=0 timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_shard_plain.json 2> gpurun_out/r02_bench_2gpu_shard_plain.err
for f in repl shard_la shard_plain; do echo "== $f"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_2gpu_$f.json").read().strip().splitlines()[-1])
    print(j["ms_per_step"], j["config"]["phases_ms"], j["config"]["predict_f_ms_M4096"], j["check"]["mll"], j["check"]["dmll"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02_bench_2gpu_$f.err").read()[-1500:])
PY
done
timeout 900 $TR --master-port 29704 profiles/tools/run_c4.py --n 49152 --d 16 --m 2048 --steps 2 > gpurun_out/r02_c4_small_2gpu.json 2> gpurun_out/r02_c4_small_2gpu.err
tail -n 3 gpurun_out/r02_c4_small_2gpu.json | cut -c1-1500; tail -n 5 gpurun_out/r02_c4_small_2gpu.err

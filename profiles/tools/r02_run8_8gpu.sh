#!/bin/bash
# round 2, 8-GPU call: C4 (N=131072, row-sharded), C2 scaling at 8 and 4 GPUs in both storage modes, 8-rank parity worker
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 0. quick 8-rank sanity of the sharded NCCL path at a small size (seconds)
timeout 240 $TR --nproc-per-node 8 --master-port 29801 profiles/tools/run_c4.py --npts 16384 --dim 16 --mtest 512 --evals 1 > gpurun_out/r02_c4_sanity_8gpu.json 2> gpurun_out/r02_c4_sanity_8gpu.err
tail -n 2 gpurun_out/r02_c4_sanity_8gpu.json | cut -c1-600
# 1. C4 at full size
if grep -q "C4_RESULT PASS" gpurun_out/r02_c4_sanity_8gpu.json; then
  timeout 600 $TR --nproc-per-node 8 --master-port 29802 profiles/tools/run_c4.py --evals 1 > gpurun_out/r02_c4_8gpu.json 2> gpurun_out/r02_c4_8gpu.err
  tail -n 2 gpurun_out/r02_c4_8gpu.json | cut -c1-2500; tail -n 3 gpurun_out/r02_c4_8gpu.err
else
  echo "sanity failed"; tail -n 20 gpurun_out/r02_c4_sanity_8gpu.err
fi
# 2. C2 strong scaling, both storage modes
for G in 8 4; do
  timeout 300 $TR --nproc-per-node $G --master-port 2981$G bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r02_bench_${G}gpu_repl.json 2> gpurun_out/r02_bench_${G}gpu_repl.err
  GPB200_P2P=0 GPB200_SHARD=1 timeout 300 $TR --nproc-per-node $G --master-port 2982$G bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r02_bench_${G}gpu_shard.json 2> gpurun_out/r02_bench_${G}gpu_shard.err
done
GPB200_P2P=0 GPB200_SHARD=1 GPB200_SHARD_RB=2 timeout 300 $TR --nproc-per-node 8 --master-port 29831 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu_shard_rb2.json 2> gpurun_out/r02_bench_8gpu_shard_rb2.err
python - <<'PY'
import json
for f in ("8gpu_repl","8gpu_shard","8gpu_shard_rb2","4gpu_repl","4gpu_shard"):
    try:
        j=json.loads(open("gpurun_out/r02_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), j["config"]["phases_ms"], round(j["config"]["predict_f_ms_M4096"],1), j["check"]["mll"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "no line", e); print(open("gpurun_out/r02_bench_%s.err"%f).read()[-800:])
PY
# 3. 8-rank parity worker (replicated p2p / nccl, sharded, distributed FITC against the oracle)
timeout 400 $TR --nproc-per-node 8 --master-port 29841 tests/mgpu_worker.py > gpurun_out/r02_mgpu_worker_8.log 2>&1
grep -E "MGPU" gpurun_out/r02_mgpu_worker_8.log | tail -12

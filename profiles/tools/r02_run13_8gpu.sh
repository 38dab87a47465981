#!/bin/bash
# round 2, second 8-GPU call: C2 strong scaling after the leaf / TRSV / compact-grid / batched-LAUUM work, 8-rank parity worker,
# C4 again
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GPB200_LEAF=${LEAF:-1}; export GPB200_LEAF
timeout 400 $TR --nproc-per-node 8 --master-port 29941 tests/mgpu_worker.py > gpurun_out/r02b_mgpu_worker_8.log 2>&1
grep -E "MGPU" gpurun_out/r02b_mgpu_worker_8.log | tail -8
for G in 8 4; do
  timeout 300 $TR --nproc-per-node $G --master-port 2991$G bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r02b_bench_${G}gpu_repl.json 2> gpurun_out/r02b_bench_${G}gpu_repl.err
done
GPB200_P2P=0 GPB200_SHARD=1 timeout 300 $TR --nproc-per-node 8 --master-port 29928 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02b_bench_8gpu_shard.json 2> gpurun_out/r02b_bench_8gpu_shard.err
python - <<'PY'
import json
for f in ("8gpu_repl","4gpu_repl","8gpu_shard"):
    try:
        j=json.loads(open("gpurun_out/r02b_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"],1), j["config"]["phases_ms"], round(j["config"]["predict_f_ms_M4096"],1), j["check"]["mll"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "no line", e); print(open("gpurun_out/r02b_bench_%s.err"%f).read()[-800:])
PY
timeout 600 $TR --nproc-per-node 8 --master-port 29942 profiles/tools/run_c4.py --evals 1 > gpurun_out/r02b_c4_8gpu.json 2> gpurun_out/r02b_c4_8gpu.err
tail -n 2 gpurun_out/r02b_c4_8gpu.json | cut -c1-1200

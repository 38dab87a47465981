#!/bin/bash
# round 2, call 1 (1 GPU): gate (all -m gpu tests, no -x), smoke, bench both arms, full-size CPU record
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_run1_gpu.txt; nproc >> gpurun_out/r02_run1_gpu.txt; free -g >> gpurun_out/r02_run1_gpu.txt
(timeout 1500 python profiles/tools/cpu_full_size.py 32768 > gpurun_out/r02_cpu_full_size.log 2>&1) &
CPUJOB=$!
timeout 2400 python -m pytest tests -q -m gpu -s --durations=15 > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_1.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke_1.log
wait $CPUJOB
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench exit $?" >> gpurun_out/r02_bench_1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_1.json 2>> gpurun_out/r02_bench_1.err
grep -E "passed|failed|error" gpurun_out/r02_pytest_gpu_1.log | tail -5
tail -n 2 gpurun_out/r02_smoke_1.log gpurun_out/r02_bench_1.err gpurun_out/r02_cpu_full_size.log
cat gpurun_out/r02_bench_1.json | cut -c1-1500

#!/bin/bash
# round 2, call 2 (1 GPU): sharded schedules as in-process groups; clean full-size CPU record (nothing CPU-heavy beside it)
mkdir -p gpurun_out
(timeout 1500 python profiles/tools/cpu_full_size.py 32768 > gpurun_out/r02_cpu_full_size.log 2>&1) &
CPUJOB=$!
timeout 1200 python -m pytest tests/test_gpu_shard.py -q -m gpu -s -x --durations=10 > gpurun_out/r02_pytest_shard_1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_shard_1.log
tail -n 40 gpurun_out/r02_pytest_shard_1.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py -q -m gpu > gpurun_out/r02_pytest_parity_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_parity_2.log
tail -n 5 gpurun_out/r02_pytest_parity_2.log
wait $CPUJOB
tail -n 2 gpurun_out/r02_cpu_full_size.log
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -s -k c5 > gpurun_out/r02_pytest_c5_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_c5_2.log
tail -n 12 gpurun_out/r02_pytest_c5_2.log

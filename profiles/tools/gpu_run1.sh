#!/bin/bash
# first GPU visit: FP64 pipe calibration, kernel-level parity, end-to-end parity, phase timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 ./profiles/tools/fp64_peak > gpurun_out/fp64_peak.txt 2>&1
timeout 300 python profiles/tools/cublas_dgemm.py > gpurun_out/cublas_dgemm.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/test_gemm.log 2>&1
echo "gemm tests exit $?" >> gpurun_out/test_gemm.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/test_parity.log 2>&1
echo "parity tests exit $?" >> gpurun_out/test_parity.log
for n in 4096 16384; do timeout 300 python profiles/tools/probe.py $n 512 0 >> gpurun_out/probe.txt 2>&1; done
timeout 300 python profiles/tools/probe.py 16384 512 1 >> gpurun_out/probe.txt 2>&1
timeout 600 python profiles/tools/probe.py 32768 512 0 4096 >> gpurun_out/probe.txt 2>&1
tail -5 gpurun_out/test_gemm.log gpurun_out/test_parity.log gpurun_out/probe.txt gpurun_out/fp64_peak.txt gpurun_out/cublas_dgemm.txt

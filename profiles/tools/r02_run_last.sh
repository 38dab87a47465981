#!/bin/bash
# last GPU slot of round 2: kernel-zoo parity of the shared-transcendental gradient leaves + C3 A/B timing
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -q -x -k "kernel_zoo or gram_matches" > gpurun_out/r02_last_zoo.log 2>&1
echo "zoo rc=$?"; tail -2 gpurun_out/r02_last_zoo.log
timeout 25 python profiles/tools/probe_c3.py > gpurun_out/r02_last_c3_new.txt 2>&1; echo "new rc=$?"; cat gpurun_out/r02_last_c3_new.txt
GPB200_PROBE_LIB=$PWD/profiles/tools/_prev/libgpb200_prev.so timeout 25 python profiles/tools/probe_c3.py > gpurun_out/r02_last_c3_prev.txt 2>&1; echo "prev rc=$?"; cat gpurun_out/r02_last_c3_prev.txt

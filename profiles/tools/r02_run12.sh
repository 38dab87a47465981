#!/bin/bash
# 1 GPU: leaf v3 (fixed) + resident-tile TRSV: parity, durations, phase timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "blocked or gram_matches or blocking or triangular or not_positive" > gpurun_out/r02_pytest_12.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_12.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_12.log | head
for L in 1; do
GPB200_LEAF=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:potrf128" -c 20 --csv --log-file gpurun_out/r02_leaf_durations_$L.csv python profiles/tools/probe.py 4096 > /dev/null 2>&1
echo "leaf variant $L:"; grep -E "potrf128" gpurun_out/r02_leaf_durations_$L.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3
done
timeout 600 python profiles/tools/probe_leaf.py 32768 > gpurun_out/r02_probe_leaf_12.txt 2>&1; cat gpurun_out/r02_probe_leaf_12.txt
for V in 2 1; do GPB200_TRSV=$V timeout 300 python profiles/tools/probe.py 32768 2>&1 | tail -1 | cut -c1-330; done

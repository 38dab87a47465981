#!/bin/bash
# round 2, call (1 GPU): blocked leaf, append, cheaper exp2 / prefetching trace: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_fullsize.py -q -m gpu -s --durations=6 > gpurun_out/r02_pytest_6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_6.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_6.log | head -30
timeout 600 python profiles/tools/probe_gram.py 32768 > gpurun_out/r02_probe_gram_6.txt 2>&1; cat gpurun_out/r02_probe_gram_6.txt
timeout 600 python profiles/tools/probe_leaf.py 32768 > gpurun_out/r02_probe_leaf_6.txt 2>&1; cat gpurun_out/r02_probe_leaf_6.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:potrf128" -c 24 --csv --log-file gpurun_out/r02_leaf_durations.csv python profiles/tools/probe_leaf.py 4096 > /dev/null 2>&1
grep -E "potrf128" gpurun_out/r02_leaf_durations.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12

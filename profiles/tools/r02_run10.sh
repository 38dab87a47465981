#!/bin/bash
# 1 GPU: blocked leaf after template-izing its serial parts: parity + per-launch durations + Cholesky phase timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "blocked or gram_matches or blocking" > gpurun_out/r02_pytest_10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_10.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_10.log | head
for L in 1 0; do
GPB200_LEAF=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:potrf128" -c 20 --csv --log-file gpurun_out/r02_leaf_durations_$L.csv python profiles/tools/probe.py 4096 > /dev/null 2>&1
echo "leaf variant $L:"; grep -E "potrf128" gpurun_out/r02_leaf_durations_$L.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3
done
timeout 600 python profiles/tools/probe_leaf.py 32768 > gpurun_out/r02_probe_leaf_10.txt 2>&1; cat gpurun_out/r02_probe_leaf_10.txt
GPB200_LEAF=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:potrf128_blk" -s 2 -c 1 -o gpurun_out/r02_prof_leaf_blk2 -f python profiles/tools/probe.py 2048 > gpurun_out/r02_ncu_leaf_blk2.log 2>&1

#!/bin/bash
# 1 GPU: compact ownership grids + batched LAUUM (group tests), source-level ncu of both diagonal-tile kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r02_pytest_9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_9.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_9.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
GPB200_LEAF=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:potrf128_blk" -s 2 -c 1 -o gpurun_out/r02_prof_leaf_blk -f python profiles/tools/probe.py 2048 > gpurun_out/r02_ncu_leaf_blk.log 2>&1
GPB200_LEAF=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:potrf128_inv" -s 2 -c 1 -o gpurun_out/r02_prof_leaf_col -f python profiles/tools/probe.py 2048 > gpurun_out/r02_ncu_leaf_col.log 2>&1
ls -la gpurun_out/r02_prof_leaf_*.ncu-rep

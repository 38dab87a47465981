#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/probe3.txt
for nb in 0 512 2048 4096; do timeout 300 python profiles/tools/probe.py 32768 $nb 0 >> gpurun_out/probe3.txt 2>&1; done
timeout 300 python profiles/tools/probe.py 16384 0 0 >> gpurun_out/probe3.txt 2>&1
timeout 300 python profiles/tools/probe.py 8192 0 0 >> gpurun_out/probe3.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potrf128 -c 4 --csv --log-file gpurun_out/leaf.csv python profiles/tools/probe.py 4096 0 0 > /dev/null 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/probe3.txt; grep potrf gpurun_out/leaf.csv | cut -d, -f5,15 | head -4

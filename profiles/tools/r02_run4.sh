#!/bin/bash
# round 2, call 4 (1 GPU): CV / rand / sparse full-cov parity, conflict-free Gram/trace kernels (timing + ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fitc.py tests/test_gpu_shard.py -q -m gpu -s --durations=6 > gpurun_out/r02_pytest_4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_4.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_4.log | head -30
timeout 600 python profiles/tools/probe_gram.py 32768 > gpurun_out/r02_probe_gram_4.txt 2>&1; cat gpurun_out/r02_probe_gram_4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gram_seiso|trace_seiso" -c 2 -o gpurun_out/r02_prof_gram_fast2 -f python profiles/tools/probe_gram.py 32768 > gpurun_out/r02_ncu_gram2.log 2>&1
ncu -i gpurun_out/r02_prof_gram_fast2.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/r02_ncu_gram_metrics2.csv 2>&1
cut -c1-1800 gpurun_out/r02_ncu_gram_metrics2.csv

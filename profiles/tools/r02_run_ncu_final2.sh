#!/bin/bash
# the three dominant launches of gpb200_dgemm_nt_tma at the end of the first step (two top inverse merges + W'W)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 521 -c 6 -o gpurun_out/r02_prof_gemm_top -f python profiles/tools/probe.py 32768 0 0 > gpurun_out/r02_ncu_gemm_top.log 2>&1
ncu -i gpurun_out/r02_prof_gemm_top.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct 2>/dev/null | cut -c1-60,300-

#!/bin/bash
# round 2: --set full captures of the final kernels (top Schur update, W'W, Gram, trace, leaf, solves)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 255 -c 1 -o gpurun_out/r02_prof_schur_top -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 526 -c 1 -o gpurun_out/r02_prof_wtw -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gram_seiso|trace_seiso|potrf128|trsv_fwd|trsv_bwd" -c 5 -o gpurun_out/r02_prof_misc_final -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
ls -la gpurun_out/r02_prof_schur_top.ncu-rep gpurun_out/r02_prof_wtw.ncu-rep gpurun_out/r02_prof_misc_final.ncu-rep

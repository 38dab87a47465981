#!/bin/bash
# final round-1 evidence on one B200: tests, smoke, bench (both arms), ncu launch list + full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 790 -c 790 --csv --log-file gpurun_out/launches_final.csv python profiles/tools/probe.py 32768 0 0 > gpurun_out/ncu_list_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 255 -c 1 -o gpurun_out/prof_schur_top -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 526 -c 1 -o gpurun_out/prof_wtw -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gram_lower|trace_kernel|potrf128|trsv_fwd" -c 4 -o gpurun_out/prof_misc_final -f python profiles/tools/probe.py 32768 0 0 > /dev/null 2>&1
tail -n 3 gpurun_out/pytest_gpu_final.log gpurun_out/smoke_final.log gpurun_out/bench_final.err
cat gpurun_out/bench_final.json

#!/bin/bash
# second GPU visit: full gpu test suite, smoke, bench, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of one full step (second repetition of the probe): cold-cache, serialised -> compare SHARES
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1680 -c 1680 --csv --log-file gpurun_out/launches.csv python profiles/tools/probe.py 32768 512 0 > gpurun_out/ncu_list.log 2>&1
# full captures: panel TRSM + trailing SYRK of outer step 14, the W'W launch, Gram, trace
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 208 -c 2 -o gpurun_out/prof_gemm -f python profiles/tools/probe.py 32768 512 0 > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_tma -s 906 -c 1 -o gpurun_out/prof_lauum -f python profiles/tools/probe.py 32768 512 0 > gpurun_out/ncu_lauum.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gram_lower|trace_kernel|potrf128" -c 3 -o gpurun_out/prof_misc -f python profiles/tools/probe.py 32768 512 0 > gpurun_out/ncu_misc.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err
cat gpurun_out/bench.json gpurun_out/bench_ref.json

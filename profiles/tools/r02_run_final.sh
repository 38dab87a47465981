#!/bin/bash
# round 2, final 1-GPU gate: every -m gpu test, smoke, bench (both arms), launch list + ncu of the step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_final.log
grep -E "passed|failed|^E  |Error" gpurun_out/r02_pytest_gpu_final.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke_final.log; tail -n 3 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?" >> gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bench_final.err
cut -c1-1200 gpurun_out/r02_bench_final.json; cut -c1-600 gpurun_out/r02_bench_ref_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 800 --csv --log-file gpurun_out/r02_launches_final.csv python profiles/tools/probe.py 32768 0 0 > gpurun_out/r02_ncu_list_final.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_final.csv")) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
names, durs = rows[hdr].index("Kernel Name"), len(rows[hdr]) - 1
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 2:]:
    try:
        tot[r[names].split("(")[0][-60:]] += float(r[durs]); cnt[r[names].split("(")[0][-60:]] += 1
    except Exception:
        pass
s = sum(tot.values())
with open("gpurun_out/r02_launch_share_final.txt", "w") as f:
    for k, v in tot.most_common(14):
        line = "%-62s %5d launches %10.3f ms %5.1f %%" % (k, cnt[k], v * 1e-6, 100 * v / s)
        print(line); f.write(line + "\n")
PY

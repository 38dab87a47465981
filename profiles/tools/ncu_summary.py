"""Extract the judged metrics from an .ncu-rep into a small text summary:
   python profiles/tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx_ncu.txt"""
import csv, subprocess, sys, io
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print('# source:', rep)
for r in rows[2:]:
    print('-' * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('%-75s %s %s' % (k, r[i], units[i]))

"""Summarise an .ncu-rep (raw page) into a small CSV for profiles/:  python scripts/ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_local_loads',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [(k, hdr.index(k)) for k in WANT if k in hdr]
with open(sys.argv[2], 'w', newline='') as fh:
    w = csv.writer(fh)
    w.writerow([k for k, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
print(open(sys.argv[2]).read())

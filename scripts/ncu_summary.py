"""Summarise an .ncu-rep (raw page) into a small CSV: python scripts/ncu_summary.py in.ncu-rep out.csv"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_register_spilling',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
with open(out, 'w') as fh:
    w = csv.writer(fh)
    w.writerow(['launch', 'metric', 'unit', 'value'])
    for li, vals in enumerate(rows[2:]):
        for k in keys:
            if k in hdr:
                i = hdr.index(k); w.writerow([li, k, units[i], vals[i]])
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
                w.writerow([li, h, units[i], vals[i]])
print(open(out).read())

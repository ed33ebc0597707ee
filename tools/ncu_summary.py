"""Prints the metrics we judge kernels by from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat and pat not in d['Kernel Name']:
        continue
    print('====', d['Kernel Name'][:90])
    for k in hdr:
        if k in KEYS or ('issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(d[k] or 0) > 0.2):
            print('   %-90s %s' % (k, d[k]))

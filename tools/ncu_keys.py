#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (run here, no GPU): python tools/ncu_keys.py file.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__cycles_elapsed.avg',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("##", d['Kernel Name'][:90], d.get('Grid Size', ''), d.get('Block Size', ''))
    for k in KEYS:
        if k in d: print('  %-80s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = sorted(((float(v.replace(',', '')), k) for k, v in d.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and v), reverse=True)
    print('  stalls/issue:', ', '.join('%s %.2f' % (k.split('issue_stalled_')[1].split('_per_')[0], v) for v, k in st[:7]))

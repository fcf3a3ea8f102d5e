"""dev: condense an `ncu --set full --csv --page raw` export to the columns DESIGN.md argues with.
usage: ncu -i capture.ncu-rep --page raw --csv > raw.csv; python tests/tools/ncu_condense.py raw.csv out.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
idx = []
for w in want:
    m = [j for j, c in enumerate(h) if c == w] or [j for j, c in enumerate(h) if w in c]
    idx.append(m[0] if m else None)
out = csv.writer(open(sys.argv[2], 'w'))
out.writerow([w + (' [' + rows[hdr + 1][j] + ']' if j is not None and rows[hdr + 1][j] else '') for w, j in zip(want, idx)])
for r in rows[hdr + 2:]:
    if len(r) < len(h):
        continue
    vals = [(r[j] if j is not None else '') for j in idx]
    vals[0] = vals[0].replace('void ', '').replace('vgtkb::', '')[:70]
    out.writerow(vals)
print("wrote", sys.argv[2])

#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with ncu -i): headline metrics, stall reasons, SASS hot segments."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size']
for d in data:
    print('--- launch:', d[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want:
            print(f"  {h} [{units[i]}] = {d[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
ia, isamp = h2.index('Instructions Executed'), h2.index('# Samples')
kern, cur = [], []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        kern.append(cur); cur = []; continue
    if r and r[0] == 'Address':
        continue
    if len(r) > ia:
        cur.append(r)
kern.append(cur)
k0 = kern[0]
tot = sum(int(r[ia]) for r in k0)
print('first launch: warp instructions', tot, 'static', len(k0))
stall = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[h2.index(h)]) for r in k0) for h in stall}
s = sum(agg.values()) or 1
print('stalls:', ', '.join(f"{h[6:]} {100*v/s:.0f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
if len(sys.argv) > 2:
    import pickle
    pickle.dump((h2, k0), open(sys.argv[2], 'wb'))

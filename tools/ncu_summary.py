#!/usr/bin/env python3
"""Summaries of gpurun_out/ ncu artefacts: launch-share table from a launch list CSV and the
headline metrics of a --set full capture.  Usage: tools/ncu_summary.py <tag>"""
import collections, csv, subprocess, sys
tag = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in data:
    if len(r) > vi:
        agg[r[ki].split("(")[0][-40:]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"# launch list ({tag}): per-launch device time, cold-cache & serialised (compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:42s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us share={100*sum(v)/tot:5.1f}%")
out = subprocess.run(["ncu", "-i", f"gpurun_out/prof_blocks_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
print(f"# ncu --set full, k_blocks ({tag})")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:70s} {rows[1][i]:>10s} " + " ".join(r[i] for r in rows[2:4]))
stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
vals = sorted(((float(rows[2][hdr.index(h)]), h) for h in stall), reverse=True)[:8]
for v, h in vals:
    print(f"  stall {h.split('stalled_')[1].replace('_per_warp_active.pct',''):30s} {v:6.1f}%")

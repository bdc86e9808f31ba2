#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals from an ncu --set full capture
(--import-source on, built with -lineinfo).  Usage: tools/ncu_lines.py <file.ncu-rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; recs = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0] not in ("", "Function Name") and r[0].isdigit():
        d = dict(zip(hdr, r))
        num = lambda v: int(v) if v and v.strip().lstrip("-").isdigit() else 0      # "-" = no data for the line
        recs.append((cur_file, int(r[0]), r[1].strip(), num(d["Instructions Executed"]), num(d["# Samples"])))
ti = sum(x[3] for x in recs); ts = sum(x[4] for x in recs)
print(f"total warp instructions {ti}, stall samples {ts}")
print("--- by instructions")
for f, ln, src, ins, smp in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{100*ins/ti:5.1f}% inst {100*smp/max(ts,1):5.1f}% smp  {f}:{ln}  {src[:90]}")
print("--- by stall samples")
for f, ln, src, ins, smp in sorted(recs, key=lambda x: -x[4])[:top]:
    print(f"{100*smp/max(ts,1):5.1f}% smp {100*ins/ti:5.1f}% inst  {f}:{ln}  {src[:90]}")

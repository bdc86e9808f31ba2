#!/usr/bin/env python3
"""SASS listing of the hottest loop of a kernel from an ncu --set full --import-source on report: executed count, stall samples, SASS.
Usage: tools/ncu_loop_listing.py <report.ncu-rep> > profiles/<round>_k_bricks_sass.txt"""
import collections, csv, subprocess, sys

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
kname = rows[0][1] if rows[0][0] == "Kernel Name" else "?"
h = rows[1]
ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
ins = [r for r in rows[2:] if len(r) > ia and r[0] not in ("Address", "Kernel Name")]
cnt = [int(r[ia]) for r in ins]; smp = [int(r[ismp]) for r in ins]
tot, stot = sum(cnt), sum(smp)
# the finest-voxel loop = the innermost backward branch whose body stores a {sdf, weight} pair of the finest level (node 72 + ...: +0x240)
import re
addr = {}
for i, r in enumerate(ins):
    try: addr[int(r[0], 16)] = i
    except ValueError: pass
lo = hi = None
for i, r in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", r[isrc])
    if not m: continue
    j = addr.get(int(m.group(1), 16))
    if j is None or j >= i: continue
    body = " ".join(x[isrc] for x in ins[j:i + 1])
    if "STG.E.64" in body and "+0x240]" in body and (lo is None or i - j < hi - lo): lo, hi = j, i
loop_exec = sum(cnt[lo:hi + 1])
iters = cnt[lo]
unroll = 2 if (hi - lo) > 450 else 1
rounds = iters * unroll
print(f"{kname[:80]}")
print(f"ncu --set full --import-source on, one launch (frame 20 of tools/prof_integrate.py, 2048^3 / 10 m, colour): {tot} warp instructions, {len(ins)} static SASS instructions, {stot} stall samples")
print(f"finest-voxel loop = SASS instructions {lo}..{hi} ({hi - lo + 1} static for {unroll} round(s), incl. the rarely taken double-precision projection): {loop_exec} executed = {100.0 * loop_exec / tot:.1f} % of the kernel,")
print(f"  ~{loop_exec / rounds:.0f} per round of 32 voxels (~{rounds} rounds), {sum(smp[lo:hi + 1])} of the stall samples")
op = collections.Counter()
for i in range(lo, hi + 1):
    s = ins[i][isrc].split()
    o = s[1] if s[0].startswith("@") else s[0]
    op[o.split(".")[0]] += cnt[i]
print("executed per round by opcode: " + ", ".join(f"{k} {v / rounds:.1f}" for k, v in op.most_common(24)))
top = sorted(range(len(ins)), key=lambda i: -smp[i])[:8]
print("instructions with the most stall samples: " + "; ".join(f"#{i} {smp[i]} [{' '.join(ins[i][isrc].split()[:3])}]" for i in top))
print()
print("index  executed  samples  SASS")
for i in range(lo, hi + 1):
    print(f"{i:5d} {cnt[i]:9d} {smp[i]:5d}  {ins[i][isrc].strip()[:110]}")

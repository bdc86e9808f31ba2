#!/usr/bin/env python3
"""Copy the judged summaries of a gpurun profile set into profiles/ (tracked).  Usage: tools/save_profiles.py <tag> <round>"""
import csv, json, os, shutil, subprocess, sys
tag, rnd = sys.argv[1], sys.argv[2]
os.makedirs("profiles", exist_ok=True)
def run(cmd): return subprocess.run(cmd, capture_output=True, text=True).stdout
open(f"profiles/{rnd}_integrate_summary.txt", "w").write(run([sys.executable, "tools/ncu_summary.py", tag]))
open(f"profiles/{rnd}_k_blocks_lines.txt", "w").write(run([sys.executable, "tools/ncu_lines.py", f"gpurun_out/prof_blocks_{tag}.ncu-rep", "30"]))
open(f"profiles/{rnd}_k_render_lines.txt", "w").write(run([sys.executable, "tools/ncu_lines.py", f"gpurun_out/prof_render_{tag}.ncu-rep", "30"]))
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{rnd}_launches.csv")
shutil.copy(f"gpurun_out/bench_{tag}.json", f"profiles/{rnd}_bench.json")
shutil.copy(f"gpurun_out/bench_ref_{tag}.json", f"profiles/{rnd}_bench_reference.json")
shutil.copy(f"gpurun_out/extra_{tag}.json", f"profiles/{rnd}_render_mesh.json")
def raw(rep, want):
    rows = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr = rows[0]; out = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w); out[w] = {"unit": rows[1][i], "values": [r[i] for r in rows[2:]]}
    return out
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
kb = raw(f"gpurun_out/prof_blocks_{tag}.ncu-rep", want); kr = raw(f"gpurun_out/prof_render_{tag}.ncu-rep", want)
json.dump({"k_blocks": kb, "k_render": kr}, open(f"profiles/{rnd}_ncu_metrics.json", "w"), indent=1)
sc = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
def traffic(k):
    r = [float(v) for v in k["dram__bytes_read.sum"]["values"]]; w = [float(v) for v in k["dram__bytes_write.sum"]["values"]]
    return sum(r) / len(r) * sc[k["dram__bytes_read.sum"]["unit"]] + sum(w) / len(w) * sc[k["dram__bytes_write.sum"]["unit"]]
json.dump({"dram_bytes_per_launch": traffic(kb), "kernel": "k_blocks", "source": f"profiles/{rnd}_ncu_metrics.json (ncu --set full, bench.py workload, {len(kb['gpu__time_duration.sum']['values'])} launches)"},
          open("profiles/traffic.json", "w"), indent=1)
t = [float(v) for v in kr["gpu__time_duration.sum"]["values"]]
print("k_blocks traffic/launch", traffic(kb), "k_render: time", t, kr["gpu__time_duration.sum"]["unit"], "traffic", traffic(kr))

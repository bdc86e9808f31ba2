#!/bin/bash
# usage: tools/gpu_final.sh <tag> — the round's closing set on one B200: gpu tests, smoke, ncu of the dominant kernel (-> traffic.json),
# launch list, both bench arms.  Everything judged is copied from gpurun_out/ into profiles/ afterwards.
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/tests_$TAG.txt; cat gpurun_out/tests_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.txt
ncu --set full --clock-control none --import-source on -k regex:k_bricks -s 20 -c 1 -o gpurun_out/prof_bricks_$TAG python tools/prof_integrate.py > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/prof_bricks_$TAG.ncu-rep gpurun_out/ncu_metrics_$TAG.json profiles/traffic.json && cp profiles/traffic.json gpurun_out/traffic_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 48 --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_integrate.py > /dev/null 2>&1
python bench.py --impl reference --steps 3 --warmup 1 2> /dev/null | tail -1 > gpurun_out/bench_ref_$TAG.json
python bench.py 2> gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read())
r = json.loads(open("gpurun_out/bench_ref_$TAG.json").read())
print("value", d["value"], "e2e", d["e2e"]["value"], "host-load", d["host_load_leg"], "roofline", d["roofline"]["frac"], d["roofline"]["batched"]["frac"], d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], "ref arm", r.get("value"))
PY

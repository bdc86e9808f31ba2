#!/bin/bash
# usage: tools/gpu_quick2.sh <tag> [pytest-args] — GPU suite (or a subset), bench line, launch list
TAG=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -8
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
cat gpurun_out/bench_$TAG.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'us/launch',d['roofline']['us_per_launch'],'launches',d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 72 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py $TAG 2>/dev/null | head -8

#!/bin/bash
# usage: tools/gpu_round3.sh <tag> — GPU suite, k_celltop_up phase timing, bench line
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/dbg_celltop.py 2>&1 | tail -40 > gpurun_out/dbg_celltop_$TAG.txt; tail -24 gpurun_out/dbg_celltop_$TAG.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
cat gpurun_out/bench_$TAG.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'us/launch',d['roofline']['us_per_launch'],'launches',d['gpu_launches'])"

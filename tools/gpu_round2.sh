#!/bin/bash
# usage: tools/gpu_round2.sh <tag> — GPU suite + bench line + front/back-end secondary measurements
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
cat gpurun_out/bench_$TAG.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'us/launch',d['roofline']['us_per_launch'],'launches',d['gpu_launches'])"
timeout 600 python tests/perf/bench_frontback.py 2>&1 | tail -3 | tee gpurun_out/frontback_$TAG.json

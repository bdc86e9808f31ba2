#!/bin/bash
# usage: tools/gpu_quick.sh <tag>  — parity tests (integrate-only subset) + bench line + launch list
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
cat gpurun_out/bench_$TAG.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],'us/launch',d['roofline']['us_per_launch'],'launches',d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 49 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1

#!/bin/bash
# usage: tools/gpu_profile.sh <tag> — the round's profile set: bench line, launch list, full ncu of k_blocks and k_render
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 50 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_blocks -s 40 -c 2 -o gpurun_out/prof_blocks_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_celltop_up -s 40 -c 2 -o gpurun_out/prof_topup_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_render_$TAG python tests/perf/bench_extra.py --frames 12 --no-cpu --only c3 > gpurun_out/extra_ncu_$TAG.log 2>&1
python tests/perf/bench_extra.py --frames 20 > gpurun_out/extra_$TAG.json 2>&1
ls gpurun_out

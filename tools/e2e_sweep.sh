#!/bin/bash
# repeats of bench.py's end-to-end leg under different host-packing settings (run on the GPU box), interleaved
out=gpurun_out/e2e_sweep.txt; : > $out
run () { env "$@" python bench.py --no-cpu-baseline --no-host-load 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']), round(d['e2e']['value']))" >> $out; }
for r in 1 2 3; do
run A=default
run B200TSDF_PACK_THREADS=24
done
sort $out

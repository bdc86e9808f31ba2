for D in 0 1; do
B200TSDF_DEBUG=$D ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 28 --csv --log-file gpurun_out/launches_exp$D.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done

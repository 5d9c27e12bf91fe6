mkdir -p gpurun_out
for wl in cfg1 cfg2 cfg2d cfg4; do
python bench.py --steps 640 --warmup 64 --workload $wl --cpu-seconds 1 > gpurun_out/bench_${wl}_s640.json 2> gpurun_out/bench_${wl}.err; tail -2 gpurun_out/bench_${wl}.err
done
python bench.py --steps 640 --warmup 64 --workload cfg4 --mode annotated --cpu-seconds 0.5 > gpurun_out/bench_cfg4_annot_s640.json 2> gpurun_out/bench_cfg4a.err; tail -2 gpurun_out/bench_cfg4a.err

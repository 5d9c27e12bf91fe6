mkdir -p gpurun_out
for mc in 1 2 4; do
python bench.py --steps 20 --warmup 5 --min-calls $mc --cpu-seconds 1 > gpurun_out/bench_cfg3_s20_mc$mc.json 2> gpurun_out/bench_mc$mc.err; tail -2 gpurun_out/bench_mc$mc.err
done
python bench.py --steps 640 --warmup 64 --cpu-seconds 1 > gpurun_out/bench_cfg3_s640.json 2> gpurun_out/bench_s640.err; tail -2 gpurun_out/bench_s640.err

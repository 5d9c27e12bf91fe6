set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg3_s20.json 2> gpurun_out/bench_cfg3_s20.err; tail -3 gpurun_out/bench_cfg3_s20.err
python bench.py --steps 640 --warmup 64 --cpu-seconds 2 > gpurun_out/bench_cfg3_s640.json 2> gpurun_out/bench_cfg3_s640.err; tail -3 gpurun_out/bench_cfg3_s640.err
python bench.py --steps 20 --warmup 5 --workload cfg2 --cpu-seconds 2 > gpurun_out/bench_cfg2_s20.json 2> gpurun_out/bench_cfg2_s20.err; tail -3 gpurun_out/bench_cfg2_s20.err

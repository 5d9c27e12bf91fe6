set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -q --tb=short -x -k "pipeline" > gpurun_out/t_pipeline.log 2>&1; tail -30 gpurun_out/t_pipeline.log
python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_pipeline.py > gpurun_out/t_rest.log 2>&1; tail -15 gpurun_out/t_rest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg3_s20.json 2> gpurun_out/bench_cfg3_s20.err; tail -5 gpurun_out/bench_cfg3_s20.err; cut -c1-1500 gpurun_out/bench_cfg3_s20.json
python bench.py --steps 20 --warmup 5 --workload cfg2 --no-tracks --cpu-seconds 3 > gpurun_out/bench_cfg2_s20.json 2> gpurun_out/bench_cfg2_s20.err; tail -5 gpurun_out/bench_cfg2_s20.err; cut -c1-1500 gpurun_out/bench_cfg2_s20.json

mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_cfg3_s20.json 2> gpurun_out/r2_ref.err; tail -2 gpurun_out/r2_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_cfg3_s20.json 2> gpurun_out/r2_cfg3_s20.err; tail -2 gpurun_out/r2_cfg3_s20.err
python bench.py --steps 640 --warmup 64 --cpu-seconds 2 > gpurun_out/r2_cfg3_s640.json 2> gpurun_out/r2_cfg3_s640.err; tail -2 gpurun_out/r2_cfg3_s640.err
python bench.py --steps 20 --warmup 5 --workload cfg2 --cpu-seconds 3 > gpurun_out/r2_cfg2_s20.json 2> gpurun_out/r2_cfg2_s20.err; tail -2 gpurun_out/r2_cfg2_s20.err
python bench.py --steps 640 --warmup 64 --workload cfg2 --cpu-seconds 2 > gpurun_out/r2_cfg2_s640.json 2> gpurun_out/r2_cfg2_s640.err; tail -2 gpurun_out/r2_cfg2_s640.err
python bench.py --steps 640 --warmup 64 --workload cfg1 --cpu-seconds 1 > gpurun_out/r2_cfg1_s640.json 2> gpurun_out/r2_cfg1.err; tail -2 gpurun_out/r2_cfg1.err
python bench.py --steps 640 --warmup 64 --workload cfg4 --cpu-seconds 1 > gpurun_out/r2_cfg4_s640.json 2> gpurun_out/r2_cfg4.err; tail -2 gpurun_out/r2_cfg4.err
python profiles/probe_dataset.py 2>&1 | tail -8

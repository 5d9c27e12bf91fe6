mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dataset.py tests/test_gpu_svar2_dataset.py -x -q -m gpu 2>&1 | tail -5
python profiles/probe_dataset.py cfg2 > gpurun_out/probe_dataset_cfg2.log 2>&1; head -1 gpurun_out/probe_dataset_cfg2.log; sed -n 2,40p gpurun_out/probe_dataset_cfg2.log | cut -c1-160
python profiles/probe_dataset.py cfg3 > gpurun_out/probe_dataset_cfg3.log 2>&1; head -1 gpurun_out/probe_dataset_cfg3.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/r2_cfg3_s20_b.json 2> gpurun_out/r2_b.err; tail -2 gpurun_out/r2_b.err; cut -c1-300 gpurun_out/r2_cfg3_s20_b.json

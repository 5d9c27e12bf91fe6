mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_open.py -x -q > gpurun_out/variants2.log 2>&1
tail -25 gpurun_out/variants2.log

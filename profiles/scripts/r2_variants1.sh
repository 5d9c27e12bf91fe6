mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_variants.py -x -q > gpurun_out/variants1.log 2>&1
tail -40 gpurun_out/variants1.log

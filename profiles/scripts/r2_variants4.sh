timeout 40 python -m pytest tests/test_gpu_variants.py -x -q 2>&1 | tail -3

mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_open.py -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python profiles/probe_variants.py 2>&1 | tail -6 | tee gpurun_out/probe_variants.txt

mkdir -p gpurun_out
SEL="test_pipelined_loader_matches_indexing or test_pipelined_loader_subset_jitter_and_fallback"
run() { name=$1; shift; echo "=== $name"; env "$@" timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=line -k "$SEL" > gpurun_out/sanitize7_$name.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|FAILED|ERROR " gpurun_out/sanitize7_$name.log | tail -6; }
run sync GVL_TEST_SYNC=1
run blocking CUDA_LAUNCH_BLOCKING=1
run split GVL_PIPE_SPLIT=1
run sync_nograph GVL_TEST_SYNC=1 GVL_PIPE_GRAPH=0
echo "=== plain (no tool), sync fixture"
GVL_TEST_SYNC=1 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=line 2>&1 | tail -2

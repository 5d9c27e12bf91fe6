for i in 1 2; do
echo "=== racecheck copy=True run $i"
CUDA_LAUNCH_BLOCKING=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=short -x -k "test_pipelined_loader_matches_indexing and True" 2>&1 | grep -E "passed|failed|hazard|RACECHECK|Error|illegal|=========.*(in|at) |\.py:[0-9]+: in|\.cu:" | head -20
done
echo "=== memcheck copy=True"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=short -x -k "test_pipelined_loader_matches_indexing and True" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*(in|at) " | head -20

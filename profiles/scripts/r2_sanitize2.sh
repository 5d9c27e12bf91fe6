mkdir -p gpurun_out
for sel in "test_fused_track_vs_oracle and 0-0.0-True" "test_getitem_fast_path_matches_general_path and bytes" "test_getitem_fast_path_tracks" "test_pipelined_loader_matches_indexing and False"; do
  echo "=== racecheck: $sel"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "$sel" 2>&1 | grep -E "passed|failed|hazard|RACECHECK|Error|illegal|Race|=========.*(in|at) " | head -12
done
echo "=== synccheck"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "pipeline or overlapping" 2>&1 | grep -E "passed|failed|SYNCCHECK|Barrier|Error|=========.*(in|at) " | head -12
echo "=== initcheck"
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "pipeline or overlapping or stress" 2>&1 | grep -E "passed|failed|INITCHECK|Uninit|Error|=========.*(in|at) " | head -12

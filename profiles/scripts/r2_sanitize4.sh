echo "=== racecheck original selection"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "pipeline or overlapping or one_bp or test_fused_track_vs_oracle" 2>&1 | grep -E "passed|failed|hazard|RACECHECK|Error|illegal|=========.*(in|at) |FAILED" | head -20
echo "=== memcheck original selection"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "pipeline or overlapping or one_bp or test_fused_track_vs_oracle" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========.*(in|at) |FAILED" | head -30

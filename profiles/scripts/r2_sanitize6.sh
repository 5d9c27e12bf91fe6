mkdir -p gpurun_out
SEL="test_pipelined_loader_matches_indexing or test_pipelined_loader_subset_jitter_and_fallback or test_tracks_and_loader_on_the_svar2_source"
echo "=== racecheck, loader tests, CUDA graphs OFF"
GVL_PIPE_GRAPH=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "$SEL" > gpurun_out/sanitize6_race_nograph.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|illegal|FAILED" gpurun_out/sanitize6_race_nograph.log | tail -8
echo "=== racecheck, loader tests, CUDA graphs ON"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "$SEL" > gpurun_out/sanitize6_race_graph.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|illegal|FAILED" gpurun_out/sanitize6_race_graph.log | tail -8
echo "=== memcheck, loader tests, CUDA graphs ON"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "$SEL" > gpurun_out/sanitize6_mem_graph.log 2>&1; echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|FAILED" gpurun_out/sanitize6_mem_graph.log | tail -6
echo "=== racecheck, whole pipeline file, graphs OFF"
GVL_PIPE_GRAPH=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_svar2_dataset.py -m gpu -q --tb=line > gpurun_out/sanitize6_race_nograph_all.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|illegal|FAILED" gpurun_out/sanitize6_race_nograph_all.log | tail -8

mkdir -p gpurun_out
echo "=== racecheck: plan kernels (goldens, oracle parity, stress subset, svar2, tracks)"
timeout 1700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "golden or parity or test_gpu_tracks or svar2 or (stress and not slow)" > gpurun_out/sanitize5_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitize5_racecheck.log | tail -5
echo "=== memcheck: everything but the > 2^31 case"
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "not large_total" > gpurun_out/sanitize5_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize5_memcheck.log | tail -4
echo "=== synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "pipeline or overlapping or golden or svar2" > gpurun_out/sanitize5_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize5_synccheck.log | tail -4
echo "=== racecheck: pipeline"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "pipeline" > gpurun_out/sanitize5_racecheck_pipe.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|illegal" gpurun_out/sanitize5_racecheck_pipe.log | tail -6

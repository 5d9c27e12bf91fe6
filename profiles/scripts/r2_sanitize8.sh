mkdir -p gpurun_out
echo "=== memcheck: everything but the > 2^31 case (final library)"
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "not large_total" > gpurun_out/sanitize8_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize8_memcheck.log | tail -3
echo "=== racecheck: svar2 + tracks + goldens (final merge / track kernels)"
timeout 1700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -k "(svar2 and not loader) or test_gpu_tracks or golden" > gpurun_out/sanitize8_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|FAILED" gpurun_out/sanitize8_racecheck.log | tail -5
echo "=== synccheck: pipeline + svar2"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -x -k "pipeline or svar2" > gpurun_out/sanitize8_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize8_synccheck.log | tail -3

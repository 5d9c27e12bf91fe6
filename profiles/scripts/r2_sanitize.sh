mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=short -x -k "stress or pipeline or test_gpu_tracks or svar2 or aux" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q --tb=short -x -k "pipeline or overlapping or one_bp or test_fused_track_vs_oracle" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitize_racecheck.log

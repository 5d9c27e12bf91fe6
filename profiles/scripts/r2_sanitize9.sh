mkdir -p gpurun_out
SEL="test_pipelined_loader_matches_indexing or test_pipelined_loader_subset_jitter_and_fallback"
run() { name=$1; shift; echo "=== racecheck $name"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 "$@" python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=line -k "$SEL" > gpurun_out/sanitize9_$name.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|FAILED" gpurun_out/sanitize9_$name.log | tail -4; }
run only_gvl --kernel-name kns=gvl
run only_torch --kernel-name-exclude kns=gvl
run no_exec_oh --kernel-name-exclude kns=hap_exec_oh
run only_exec_oh --kernel-name kns=hap_exec_oh
run only_plan --kernel-name kns=hap_plan_par
run only_prep --kernel-name kns=batch_prep
run only_hap_exec --kernel-name kns=hap_exec_kernel

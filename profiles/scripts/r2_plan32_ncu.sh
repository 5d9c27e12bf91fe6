mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_plan_par_kernel|batch_prep' -s 6 -c 4 -o gpurun_out/prof_r2_plan32 -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg4 > gpurun_out/prof_r2_plan32.out 2>&1
tail -1 gpurun_out/prof_r2_plan32.out | cut -c1-100

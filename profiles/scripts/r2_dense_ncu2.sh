mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_exec_oh_kernel' -s 3 -c 1 -o gpurun_out/prof_r2_dense2 -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg2d > gpurun_out/prof_r2_dense2.out 2>&1
tail -1 gpurun_out/prof_r2_dense2.out | cut -c1-100

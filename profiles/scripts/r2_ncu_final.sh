mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_exec_oh_kernel|hap_plan_par_kernel|batch_prep|trk_exec3_kernel|trk_tile_prep|trk_tile_scan|track_lengths' -s 4 -c 16 -o gpurun_out/prof_r2_final -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/prof_r2_final.out 2>&1
tail -1 gpurun_out/prof_r2_final.out | cut -c1-100

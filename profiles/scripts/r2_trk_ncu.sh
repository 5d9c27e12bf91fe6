mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trk_exec3_kernel|trk_tile_prep' -s 4 -c 2 -o gpurun_out/prof_r2_trk -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/prof_r2_trk.out 2>&1
tail -2 gpurun_out/prof_r2_trk.out | cut -c1-200

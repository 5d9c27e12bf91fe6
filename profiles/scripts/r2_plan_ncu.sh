mkdir -p gpurun_out
python profiles/probe_dataset.py cfg3 2>&1 | head -2 > gpurun_out/probe_dataset_cfg3.log; cat gpurun_out/probe_dataset_cfg3.log
python profiles/probe_dataset.py cfg2 2>&1 | head -2 > gpurun_out/probe_dataset_cfg2.log; cat gpurun_out/probe_dataset_cfg2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_plan_par_kernel|trk_exec3_kernel|trk_tile_prep_kernel|batch_prep_kernel|svar2_merge' -s 8 -c 10 -o gpurun_out/prof_r2_plan -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/prof_r2_plan.out 2>&1
tail -3 gpurun_out/prof_r2_plan.out | cut -c1-300
python bench.py --steps 20 --warmup 5 --workload cfg2d --cpu-seconds 1 > gpurun_out/r2_cfg2d_s20.json 2> gpurun_out/r2_cfg2d.err; tail -2 gpurun_out/r2_cfg2d.err; cut -c1-600 gpurun_out/r2_cfg2d_s20.json
python bench.py --steps 640 --warmup 64 --workload cfg2d --cpu-seconds 1 > gpurun_out/r2_cfg2d_s640.json 2> gpurun_out/r2_cfg2d.err; tail -2 gpurun_out/r2_cfg2d.err; cut -c1-600 gpurun_out/r2_cfg2d_s640.json

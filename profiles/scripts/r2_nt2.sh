mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']))
PY
}
for nt in 256 512; do
for wl in cfg3 cfg2d; do
GVL_PLAN_NT=$nt python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload $wl > gpurun_out/nt2_${wl}_$nt.json 2>gpurun_out/ab.err; pick gpurun_out/nt2_${wl}_$nt.json
done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_plan_par_kernel' -s 4 -c 2 -o gpurun_out/prof_r2_plan2 -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/prof_r2_plan2.out 2>&1
tail -1 gpurun_out/prof_r2_plan2.out | cut -c1-100

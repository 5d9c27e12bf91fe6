mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'])
PY
}
for t in 1024 2048 4096 8192 16384; do
GVL_OH_TILE=$t python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload cfg2d > gpurun_out/tile_cfg2d_$t.json 2>gpurun_out/ab.err; pick gpurun_out/tile_cfg2d_$t.json
done
for t in 4096 8192; do
GVL_OH_TILE=$t python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/tile_cfg3_$t.json 2>gpurun_out/ab.err; pick gpurun_out/tile_cfg3_$t.json
done

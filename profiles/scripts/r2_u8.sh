mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'], d['roofline']['kernel'][:40])
PY
}
for wl in cfg3 cfg2 cfg4; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl --mode u8 > gpurun_out/u8_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/u8_${wl}.json
done
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg3 --mode annotated > gpurun_out/ann_cfg3.json 2>gpurun_out/ab.err; pick gpurun_out/ann_cfg3.json

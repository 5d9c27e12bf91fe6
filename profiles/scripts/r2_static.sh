mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'], 'trk', (d.get('tracks') or {}).get('whole_step_frac'))
PY
}
for wl in cfg4 cfg3 cfg1 cfg2 cfg2d; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/st_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/st_${wl}.json
done
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.2 --workload cfg4 > gpurun_out/st_cfg4_640.json 2>gpurun_out/ab.err; pick gpurun_out/st_cfg4_640.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.2 --workload cfg1 > gpurun_out/st_cfg1_640.json 2>gpurun_out/ab.err; pick gpurun_out/st_cfg1_640.json

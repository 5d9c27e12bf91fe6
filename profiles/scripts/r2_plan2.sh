mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], d['timed_block_ms']['min'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'api %.4g'%d['api']['value'], 'trk', d.get('tracks',{}).get('ms_per_step'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 > gpurun_out/p2_cfg3.json 2>gpurun_out/ab.err; pick gpurun_out/p2_cfg3.json
for wl in cfg2 cfg2d cfg4 cfg1; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload $wl > gpurun_out/p2_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/p2_${wl}.json
done
python bench.py --steps 640 --warmup 64 --cpu-seconds 0.5 > gpurun_out/p2_cfg3_640.json 2>gpurun_out/ab.err; pick gpurun_out/p2_cfg3_640.json

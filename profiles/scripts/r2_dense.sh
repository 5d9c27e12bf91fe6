mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'], 'trk', d.get('tracks',{}).get('ms_per_step'), d.get('tracks',{}).get('whole_step_frac'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload cfg2d > gpurun_out/d_cfg2d_20.json 2>gpurun_out/ab.err; pick gpurun_out/d_cfg2d_20.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.5 --workload cfg2d > gpurun_out/d_cfg2d_640.json 2>gpurun_out/ab.err; pick gpurun_out/d_cfg2d_640.json
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 > gpurun_out/d_cfg3_20.json 2>gpurun_out/ab.err; pick gpurun_out/d_cfg3_20.json
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload cfg4 > gpurun_out/d_cfg4_20.json 2>gpurun_out/ab.err; pick gpurun_out/d_cfg4_20.json

mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dataset.py tests/test_gpu_svar2_dataset.py -x -q -m gpu 2>&1 | tail -3
GVL_PIPE_SPLIT=1 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], d['timed_block_ms']['min'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'api %.4g'%d['api']['value'], 'trk', d.get('tracks',{}).get('ms_per_step'), d.get('tracks',{}).get('whole_step_frac'))
PY
}
for s in 20 640; do
python bench.py --steps $s --warmup 5 --cpu-seconds 0.5 > gpurun_out/f_cfg3_$s.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg3_$s.json
GVL_PIPE_SPLIT=1 python bench.py --steps $s --warmup 5 --cpu-seconds 0.5 > gpurun_out/f_cfg3_${s}_split.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg3_${s}_split.json
done
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload cfg2 > gpurun_out/f_cfg2_20.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg2_20.json
GVL_PIPE_SPLIT=1 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload cfg2 > gpurun_out/f_cfg2_20_split.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg2_20_split.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.5 --workload cfg2 > gpurun_out/f_cfg2_640.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg2_640.json
GVL_PIPE_SPLIT=1 python bench.py --steps 640 --warmup 5 --cpu-seconds 0.5 --workload cfg2 > gpurun_out/f_cfg2_640_split.json 2>gpurun_out/ab.err; pick gpurun_out/f_cfg2_640_split.json

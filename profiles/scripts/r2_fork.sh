mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dataset.py tests/test_gpu_svar2_dataset.py -x -q -m gpu 2>&1 | tail -3
GVL_PIPE_GRAPH=0 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], 'whole %.3f'%d['whole_step_frac'], 'api %.4g'%d['api']['value'], 'trk', d.get('tracks',{}).get('ms_per_step'), d.get('tracks',{}).get('whole_step_frac'), 'launches', d.get('tracks',{}).get('gpu_launches'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/fk_cfg3_20.json 2>gpurun_out/ab.err; pick gpurun_out/fk_cfg3_20.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 > gpurun_out/fk_cfg3_640.json 2>gpurun_out/ab.err; pick gpurun_out/fk_cfg3_640.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 --workload cfg1 > gpurun_out/fk_cfg1_640.json 2>gpurun_out/ab.err; pick gpurun_out/fk_cfg1_640.json
python profiles/probe_dataset.py cfg2 2>&1 | head -1

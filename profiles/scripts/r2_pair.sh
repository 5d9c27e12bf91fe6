mkdir -p gpurun_out
python -m pytest tests/test_gpu_tracks.py tests/test_gpu_stress.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'trk', d.get('tracks',{}).get('ms_per_step'), d.get('tracks',{}).get('whole_step_frac'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/pr_5.json 2>gpurun_out/ab.err; pick gpurun_out/pr_5.json
GVL_LIB_NAME=libgvl_pair4.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/pr_4.json 2>gpurun_out/ab.err; pick gpurun_out/pr_4.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 > gpurun_out/pr_5_640.json 2>gpurun_out/ab.err; pick gpurun_out/pr_5_640.json
python profiles/probe_tracks.py 2>&1 | tail -1
GVL_LIB_NAME=libgvl_pair4.so python profiles/probe_tracks.py 2>&1 | tail -1

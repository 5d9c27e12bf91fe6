mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'trk', d.get('tracks',{}).get('ms_per_step'), d.get('tracks',{}).get('whole_step_frac'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/mb_4.json 2>gpurun_out/ab.err; pick gpurun_out/mb_4.json
GVL_LIB_NAME=libgvl_minb5.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/mb_5.json 2>gpurun_out/ab.err; pick gpurun_out/mb_5.json
python profiles/probe_tracks.py 2>&1 | tail -1
GVL_LIB_NAME=libgvl_minb5.so python profiles/probe_tracks.py 2>&1 | tail -1

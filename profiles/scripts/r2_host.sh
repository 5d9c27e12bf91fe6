mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 > gpurun_out/hl_cfg3.json 2>gpurun_out/hl.err; tail -2 gpurun_out/hl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/hl_cfg3.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'e2e_loader', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e_loader'].items() if k!='path'})
PY

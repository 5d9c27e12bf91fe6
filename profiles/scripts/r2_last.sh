python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/last_cfg3_s20.json 2> gpurun_out/last.err; tail -1 gpurun_out/last.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/last_ref.json 2>> gpurun_out/last.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/last_cfg3_s20.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/last_ref.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'e2e %.4g'%d['e2e']['value'], 'loader %.4g'%d['e2e_loader']['value'], 'trk', d['tracks']['whole_step_frac'], 'ref %.4g'%r['value'], 'launches', d['gpu_launches'])
PY

mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python profiles/probe_variants.py 2>&1 | tail -8 | tee gpurun_out/probe_variants.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/final3_cfg3_s20.json 2> gpurun_out/final3.err; tail -1 gpurun_out/final3.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final3_ref.json 2>> gpurun_out/final3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final3_cfg3_s20.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/final3_ref.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'e2e %.4g'%d['e2e']['value'], 'loader %.4g'%d['e2e_loader']['value'], 'trk', d['tracks']['whole_step_frac'], 'ref %.4g'%r['value'], 'launches', d['gpu_launches'])
PY

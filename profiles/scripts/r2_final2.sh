mkdir -p gpurun_out/final
F=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > $F/ref_cfg3_s20.json 2> $F/ref.err; tail -1 $F/ref.err
python bench.py --steps 20 --warmup 5 > $F/cfg3_s20.json 2> $F/cfg3_s20.err; tail -1 $F/cfg3_s20.err
python bench.py --steps 640 --warmup 64 --cpu-seconds 2 > $F/cfg3_s640.json 2> $F/cfg3_s640.err
python bench.py --steps 20 --warmup 5 --workload cfg2 --cpu-seconds 3 > $F/cfg2_s20.json 2> $F/cfg2_s20.err
python bench.py --steps 640 --warmup 64 --workload cfg2 --cpu-seconds 2 > $F/cfg2_s640.json 2> $F/cfg2_s640.err
python bench.py --impl reference --steps 20 --warmup 5 --workload cfg2 > $F/ref_cfg2_s20.json 2> $F/ref2.err
for wl in cfg1 cfg4 cfg2d; do
python bench.py --steps 640 --warmup 64 --workload $wl --cpu-seconds 1 > $F/${wl}_s640.json 2> $F/$wl.err
python bench.py --steps 20 --warmup 5 --workload $wl --cpu-seconds 1 > $F/${wl}_s20.json 2> $F/$wl.err
done
python bench.py --steps 640 --warmup 64 --workload cfg4 --mode annotated --cpu-seconds 1 > $F/cfg4ann_s640.json 2> $F/cfg4ann.err
python profiles/probe_dataset.py cfg2 2>&1 | head -1 > $F/probe_dataset.log; python profiles/probe_dataset.py cfg3 2>&1 | head -1 >> $F/probe_dataset.log; cat $F/probe_dataset.log
python profiles/probe_tracks.py 2>&1 | tail -3 > $F/probe_tracks.log; cat $F/probe_tracks.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f.split('/')[-1], 'value %.4g'%d['value'], 'ms/step %.5f'%d['ms_per_step'], 'whole', round(d.get('whole_step_frac',0),3), 'roof', round(r.get('frac',0),3), 'exec', r.get('launch_ms'), 'plan', r.get('plan_kernel_ms'), 'api %.4g'%(d.get('api') or {}).get('value',0), 'e2e %.4g'%d['e2e']['value'], 'cpu %.4g'%(d.get('cpu_baseline') or {}).get('value',0), 'trk', (d.get('tracks') or {}).get('ms_per_step'), (d.get('tracks') or {}).get('whole_step_frac'))
    except Exception as e: print(f, 'ERR', e)
PY

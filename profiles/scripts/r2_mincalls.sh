mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'ms/step %.5f'%d['ms_per_step'], 'whole %.3f'%d['whole_step_frac'], 'calls', d['pipeline']['device_calls_per_timed_block'], 'ring', d['pipeline']['batches_per_device_call'], 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'trk', d.get('tracks',{}).get('whole_step_frac'))
PY
}
for mc in 1 2 4; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --min-calls $mc > gpurun_out/mc_$mc.json 2>gpurun_out/mc.err; pick gpurun_out/mc_$mc.json
done

mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'trk', (d.get('tracks') or {}).get('whole_step_frac'))
PY
}
python -m pytest tests/test_gpu_svar2_dataset.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
for wl in cfg2 cfg3; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload $wl > gpurun_out/h_${wl}_20.json 2>gpurun_out/ab.err; pick gpurun_out/h_${wl}_20.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 --workload $wl > gpurun_out/h_${wl}_640.json 2>gpurun_out/ab.err; pick gpurun_out/h_${wl}_640.json
done

mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_stress.py tests/test_gpu_onehot_packed.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f'%d['roofline']['launch_ms'], 'roof %.3f'%d['roofline']['frac'])
PY
}
for wl in cfg2d cfg4 cfg1; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/u3_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/u3_${wl}.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/u3_${wl}_640.json 2>gpurun_out/ab.err; pick gpurun_out/u3_${wl}_640.json
done

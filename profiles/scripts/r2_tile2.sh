mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'])
PY
}
GVL_LIB_NAME=libgvl_tile32k.so python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_onehot_packed.py -x -q -m gpu 2>&1 | tail -2
for wl in cfg3 cfg2 cfg2d cfg1 cfg4; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload $wl > gpurun_out/t2_${wl}_16k.json 2>gpurun_out/ab.err; pick gpurun_out/t2_${wl}_16k.json
GVL_LIB_NAME=libgvl_tile32k.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload $wl > gpurun_out/t2_${wl}_32k.json 2>gpurun_out/ab.err; pick gpurun_out/t2_${wl}_32k.json
done
GVL_LIB_NAME=libgvl_tile32k.so python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 > gpurun_out/t2_cfg3_640_32k.json 2>gpurun_out/ab.err; pick gpurun_out/t2_cfg3_640_32k.json

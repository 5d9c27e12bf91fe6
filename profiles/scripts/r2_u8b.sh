mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f'%d['roofline']['launch_ms'], 'roof %.3f'%d['roofline']['frac'])
PY
}
for lib in libgvl_mu32.so libgvl_mu64.so; do
GVL_LIB_NAME=$lib python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_stress.py -x -q -m gpu 2>&1 | tail -1
for wl in cfg3 cfg2; do
GVL_LIB_NAME=$lib python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl --mode u8 > gpurun_out/u8b_${wl}_$lib.json 2>gpurun_out/ab.err; pick gpurun_out/u8b_${wl}_$lib.json
done
GVL_LIB_NAME=$lib python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg3 --mode annotated > gpurun_out/annb_cfg3_$lib.json 2>gpurun_out/ab.err; pick gpurun_out/annb_cfg3_$lib.json
done

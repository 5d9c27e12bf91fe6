mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f'%d['roofline']['launch_ms'], 'roof %.3f'%d['roofline']['frac'])
PY
}
for lib in libgvl_c6.so libgvl_c10.so; do
for wl in cfg3 cfg2d cfg1 cfg2; do
GVL_LIB_NAME=$lib python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/ct_${wl}_$lib.json 2>gpurun_out/ab.err; pick gpurun_out/ct_${wl}_$lib.json
done; done

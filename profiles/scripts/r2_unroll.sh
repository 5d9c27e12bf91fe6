mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f'%d['roofline']['launch_ms'], 'roof %.3f'%d['roofline']['frac'])
PY
}
for wl in cfg2d cfg4 cfg3; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/un_${wl}_4.json 2>gpurun_out/ab.err; pick gpurun_out/un_${wl}_4.json
for u in 1 2; do
GVL_LIB_NAME=libgvl_unroll$u.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/un_${wl}_$u.json 2>gpurun_out/ab.err; pick gpurun_out/un_${wl}_$u.json
done; done

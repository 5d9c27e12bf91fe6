mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'exec %.4f'%d['roofline']['launch_ms'], 'roof %.3f'%d['roofline']['frac'])
PY
}
for wl in cfg3 cfg4 cfg2d cfg1; do
GVL_LIB_NAME=libgvl_exp4.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/exp4_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/exp4_${wl}.json
done

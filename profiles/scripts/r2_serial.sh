mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'blk', d['timed_block_ms']['median'])
PY
}
for wl in cfg4 cfg1; do
GVL_PLAN=s python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload $wl > gpurun_out/ser_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/ser_${wl}.json
done

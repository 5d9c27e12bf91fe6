mkdir -p gpurun_out
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'])
PY
}
for nt in 32 256; do
GVL_PLAN_NT=$nt python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload cfg2 > gpurun_out/nt_cfg2_$nt.json 2>gpurun_out/ab.err; pick gpurun_out/nt_cfg2_$nt.json
GVL_PLAN_NT=$nt python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload cfg1 > gpurun_out/nt_cfg1_$nt.json 2>gpurun_out/ab.err; pick gpurun_out/nt_cfg1_$nt.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 300 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg2 > gpurun_out/launches_cfg2.out 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg2.csv')) if len(r)>10 and r[0].isdigit()]
seen=0
for r in rows:
    name=r[4].split('(')[0][-44:]; v=float(r[-1]); u=r[-2]
    if u=='ms': v*=1000
    if u=='ns': v/=1000
    if ('merge' in name or 'plan' in name or 'exec' in name or 'prep' in name) and v>10:
        seen+=1
        if seen<14: print(f"{r[0]:>5s} {name:46s} {v:9.1f} us")
PY

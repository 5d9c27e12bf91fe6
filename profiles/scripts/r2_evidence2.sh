mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
# steady-state DRAM traffic of the execute kernels: consecutive launches inside the timed blocks, no cache flush between them
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:'hap_exec_oh_kernel|trk_exec3_kernel' -s 12 -c 14 --csv --log-file gpurun_out/ncu_traffic_cfg3.csv \
    python bench.py --steps 64 --warmup 8 --cpu-seconds 0.2 --ring 16 > gpurun_out/ncu_traffic_cfg3.out 2>&1
tail -16 gpurun_out/ncu_traffic_cfg3.csv | cut -c1-200
# launch list of the default bench (shares of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/launches_r2.out 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r2.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1]); u=r[-2]
    if u=='ms': v*=1000
    if u=='ns': v/=1000
    agg[name][0]+=1; agg[name][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{k[-60:]:60s} {n:5d} {t:10.1f} us {t/n:9.2f} us/launch")
PY
python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/ev_cfg3_20.json 2>gpurun_out/ab.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ev_cfg3_20.json').read().strip().splitlines()[-1])
print('value %.4g'%d['value'], 'blk', d['timed_block_ms'], 'whole %.3f'%d['whole_step_frac'], 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'e2e %.4g'%d['e2e']['value'], 'trk', d['tracks']['ms_per_step'], d['tracks']['whole_step_frac'])
PY

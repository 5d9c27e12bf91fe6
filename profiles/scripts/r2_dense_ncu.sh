mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hap_exec_oh_kernel|hap_plan_par_kernel' -s 6 -c 4 -o gpurun_out/prof_r2_dense -f \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 --workload cfg2d > gpurun_out/prof_r2_dense.out 2>&1
tail -2 gpurun_out/prof_r2_dense.out | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/launches_r2b.out 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r2b.csv')) if len(r)>10 and r[0].isdigit()]
print(len(rows))
seen=0
for r in rows:
    name=r[4].split('(')[0][-44:]; v=float(r[-1]); u=r[-2]
    if u=='ms': v*=1000
    if u=='ns': v/=1000
    if 'trk' in name or 'tile' in name or 'scan' in name or 'ELb1' in name or ', 1>' in name or 'true' in name:
        seen+=1
        if seen<40: print(f"{r[0]:>5s} {name:46s} {v:9.1f} us")
PY

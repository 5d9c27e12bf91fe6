mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dataset.py tests/test_gpu_svar2_dataset.py tests/test_gpu_open.py tests/test_gpu_tracks.py -x -q -m gpu 2>&1 | tail -3
python profiles/probe_dataset.py cfg2 > gpurun_out/probe_dataset_cfg2.log 2>&1; head -1 gpurun_out/probe_dataset_cfg2.log; sed -n 2,16p gpurun_out/probe_dataset_cfg2.log | cut -c1-150
python profiles/probe_dataset.py cfg3 > gpurun_out/probe_dataset_cfg3.log 2>&1; head -1 gpurun_out/probe_dataset_cfg3.log
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], d['timed_block_ms']['min'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'api %.4g'%d['api']['value'], 'trk', d.get('tracks',{}).get('ms_per_step'))
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 > gpurun_out/ab_default.json 2>gpurun_out/ab.err; pick gpurun_out/ab_default.json
GVL_LIB_NAME=libgvl_occ1280.so python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 > gpurun_out/ab_occ.json 2>gpurun_out/ab.err; pick gpurun_out/ab_occ.json
for wl in cfg2d cfg2 cfg4; do
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload $wl > gpurun_out/ab_${wl}.json 2>gpurun_out/ab.err; pick gpurun_out/ab_${wl}.json
GVL_PLAN_NT=512 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload $wl > gpurun_out/ab_${wl}_512.json 2>gpurun_out/ab.err; pick gpurun_out/ab_${wl}_512.json
GVL_PLAN_NT=256 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --workload $wl > gpurun_out/ab_${wl}_256.json 2>gpurun_out/ab.err; pick gpurun_out/ab_${wl}_256.json
done

mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "svar2 or stress" 2>&1 | tail -3
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'blk', d['timed_block_ms']['median'], 'whole %.3f'%d['whole_step_frac'], 'exec %.4f plan %.4f'%(d['roofline']['launch_ms'], d['roofline']['plan_kernel_ms']), 'roof %.3f'%d['roofline']['frac'], 'api %.4g'%d['api']['value'], 'e2e %.4g'%d['e2e']['value'])
PY
}
python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 --workload cfg2 > gpurun_out/m_cfg2_20.json 2>gpurun_out/ab.err; pick gpurun_out/m_cfg2_20.json
python bench.py --steps 640 --warmup 5 --cpu-seconds 0.3 --workload cfg2 > gpurun_out/m_cfg2_640.json 2>gpurun_out/ab.err; pick gpurun_out/m_cfg2_640.json

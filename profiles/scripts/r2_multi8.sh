mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; tail -2 gpurun_out/bench_cfg3_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 640 --warmup 64 --cpu-seconds 1 > gpurun_out/bench_cfg3_n${N}_s640.json 2> gpurun_out/bench_cfg3_n${N}_s640.err; tail -2 gpurun_out/bench_cfg3_n${N}_s640.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_cfg3_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n', d['n_gpus'], 'steps', d['steps'], 'value %.4g'%d['value'], 'ms/step %.5f'%d['ms_per_step'], 'whole %.3f'%d['whole_step_frac'], 'e2e %.4g'%d['e2e']['value'], 'gather', json.dumps(d.get('gather'))[:300], 'per_rank', [round(x,4) for x in d['per_rank_ms']])
    except Exception as e: print(f, 'ERR', e)
PY

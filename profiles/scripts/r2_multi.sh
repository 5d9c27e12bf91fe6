mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; tail -3 gpurun_out/bench_cfg3_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 20 --warmup 5 > gpurun_out/bench_cfg3_ref_n$N.json 2> gpurun_out/bench_cfg3_ref_n$N.err; tail -3 gpurun_out/bench_cfg3_ref_n$N.err

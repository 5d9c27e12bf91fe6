set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "tracks or dataset or pipeline or open or large_total or stress or fullsize" > gpurun_out/t_tracks.log 2>&1; tail -15 gpurun_out/t_tracks.log
python profiles/probe_tracks.py 2>&1 | tail -2
PROBE_VKB=0 python profiles/probe_tracks.py 2>&1 | tail -1
PROBE_RING=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trk_exec3 -s 2 -c 1 -o gpurun_out/prof_trk3 python profiles/probe_tracks.py > gpurun_out/ncu_trk3.log 2>&1; tail -2 gpurun_out/ncu_trk3.log

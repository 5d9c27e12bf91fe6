set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "tracks or dataset or pipeline or open or large_total or stress or fullsize" > gpurun_out/t_tracks.log 2>&1; tail -15 gpurun_out/t_tracks.log
python profiles/probe_tracks.py 2>&1 | tail -1
PROBE_VKB=0 python profiles/probe_tracks.py 2>&1 | tail -1
PROBE_RING=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:trk_ -s 12 -c 8 python profiles/probe_tracks.py 2>&1 | grep -E "trk_|duration|inst_executed" | head -40

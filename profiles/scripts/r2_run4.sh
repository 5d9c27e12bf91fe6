set -x
mkdir -p gpurun_out
python profiles/probe_tracks.py 2>&1 | tail -3
PROBE_VKB=0 python profiles/probe_tracks.py 2>&1 | tail -2
PROBE_RING=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trk_exec2 -s 2 -c 1 -o gpurun_out/prof_trk2 python profiles/probe_tracks.py > gpurun_out/ncu_trk2.log 2>&1; tail -3 gpurun_out/ncu_trk2.log

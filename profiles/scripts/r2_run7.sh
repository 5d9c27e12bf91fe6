mkdir -p gpurun_out
PROBE_RING=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trk_exec3 -s 2 -c 1 -o gpurun_out/prof_trk3b python profiles/probe_tracks.py > gpurun_out/ncu_trk3b.log 2>&1; tail -2 gpurun_out/ncu_trk3b.log

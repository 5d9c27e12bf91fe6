set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "tracks or dataset or pipeline or open or large_total or stress or svar2 or aux" > gpurun_out/t_tracks.log 2>&1; tail -40 gpurun_out/t_tracks.log

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "svar2 or pipeline" > gpurun_out/t_svar2.log 2>&1; tail -25 gpurun_out/t_svar2.log

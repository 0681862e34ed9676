# round 2, run AO: GPU suite after the schedule re-tune
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -6 gpurun_out/t_gpu.log

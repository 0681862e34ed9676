# round 2, run AA: GCN inference kernels against the reference module's outputs; whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -15 gpurun_out/t_gpu.log

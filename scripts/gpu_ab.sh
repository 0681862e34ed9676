# GPU suite on the in-tree build, then a same-box A/B of variant builds (scripts/ab_rounds.py).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -5 gpurun_out/t_gpu.log
timeout 600 python scripts/ab_rounds.py $AB_LIBS > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log

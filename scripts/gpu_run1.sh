mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -k "tensor_core_scores" -x -q > gpurun_out/t1.log 2>&1
rc=$?; echo "t1 rc=$rc" >> gpurun_out/t1.log
tail -30 gpurun_out/t1.log
if [ $rc -eq 0 ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2.log
  tail -60 gpurun_out/t2.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
  tail -5 gpurun_out/bench.log
fi

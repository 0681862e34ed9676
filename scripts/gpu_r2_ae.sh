# round 2, run AE: compose / gather kernels specialised per width -- parity tests + the bench line (aux_kernels block)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/bench.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['e2e']['value']); print(json.dumps(j['aux_kernels'])[:700])"

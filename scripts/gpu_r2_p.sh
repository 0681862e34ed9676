# round 2, run P (gpurun --gpus N): the C4 scaling line only
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_p2p.log 2>&1; echo "c4 rc=$?"
grep '^{' gpurun_out/bench_n${N}_p2p.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_p2p.log | head -5

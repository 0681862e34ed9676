# Bench only (no pytest) at N GPUs: peer-store exchange vs NCCL at B=4096, then B=64 and B=1 with the peer exchange.
N=${1:-8}
mkdir -p gpurun_out
for EX in p2p nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --exchange $EX > gpurun_out/bench_n${N}_$EX.log 2>&1
  echo "bench N=$N $EX rc=$?"; grep '^{' gpurun_out/bench_n${N}_$EX.log | cut -c1-260; grep -i "error\|Traceback" gpurun_out/bench_n${N}_$EX.log | head -5
done
for B in 64 1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 50 --warmup 5 --batch $B > gpurun_out/bench_n${N}_b$B.log 2>&1
  echo "bench N=$N B=$B rc=$?"; grep '^{' gpurun_out/bench_n${N}_b$B.log | cut -c1-260
done

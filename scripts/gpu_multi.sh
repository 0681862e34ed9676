# N-GPU validation (run under `gpurun --gpus N`): full parity suite incl. the 2-GPU peer-exchange test, then
# bench.py at N GPUs with the peer-store exchange and with the NCCL all-gather for comparison.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpu or peer_exchange or merge_and_shard" > gpurun_out/t2_multi.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2_multi.log
tail -5 gpurun_out/t2_multi.log
for EX in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --exchange $EX > gpurun_out/bench_n${N}_$EX.log 2>&1
  echo "bench N=$N $EX rc=$?"; grep '^{' gpurun_out/bench_n${N}_$EX.log | cut -c1-400
done
for B in 64 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 50 --warmup 5 --batch $B > gpurun_out/bench_n${N}_b$B.log 2>&1
  echo "bench N=$N B=$B rc=$?"; grep '^{' gpurun_out/bench_n${N}_b$B.log | cut -c1-330
done

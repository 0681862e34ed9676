# round 2, run K (gpurun --gpus N): owned-result exchange + fused set-up launch validated on N GPUs, schedule sweep
N=${1:-2}
GRID=${2:-0:0,4096:16,2048:32,1024:32}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpu or peer_exchange or merge_and_shard or exact_matches_oracle" > gpurun_out/t2_multi.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2_multi.log
tail -3 gpurun_out/t2_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tune_schedule_sharded.py --batches 4096,64 --grid $GRID > gpurun_out/tune_n$N.log 2>&1; echo "tune rc=$?"
grep -v RESULT gpurun_out/tune_n$N.log | grep "^B" | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_p2p.log 2>&1; echo "c4 rc=$?"
grep '^{' gpurun_out/bench_n${N}_p2p.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_p2p.log | head -5

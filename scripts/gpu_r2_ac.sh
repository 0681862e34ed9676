# round 2, run AC (gpurun --gpus N): scaling lines of HEAD -- C4 at N GPUs (+ the 2-GPU pytest at N = 2; + C5 and C3 at N = 8)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpu or peer_exchange or merge_and_shard" > gpurun_out/t2_multi.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2_multi.log
tail -3 gpurun_out/t2_multi.log
fi
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 20 --warmup 3" > gpurun_out/bench_n${N}_p2p.log 2>&1; echo "c4 rc=$?"
grep '^{' gpurun_out/bench_n${N}_p2p.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_p2p.log | head -5
if [ "$N" = "8" ]; then
timeout 900 bash -c "$(declare -f run); N=$N; run --workload c5 --steps 5 --warmup 3" > gpurun_out/bench_n${N}_c5.log 2>&1; echo "c5 rc=$?"
grep '^{' gpurun_out/bench_n${N}_c5.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_c5.log | head -5
timeout 600 bash -c "$(declare -f run); N=$N; run --workload c3 --steps 5 --warmup 3" > gpurun_out/bench_n${N}_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_n${N}_c3.log | cut -c1-300
fi

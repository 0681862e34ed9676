# round 2, run I (gpurun --gpus N): the scaling lines only -- C4 (B=4096, B=1/64 under "sweep"), C5, C3 at N GPUs.
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 20 --warmup 3" > gpurun_out/bench_n${N}_p2p.log 2>&1; echo "c4 rc=$?"
grep '^{' gpurun_out/bench_n${N}_p2p.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_p2p.log | head -5
timeout 900 bash -c "$(declare -f run); N=$N; run --workload c5 --steps 5 --warmup 3" > gpurun_out/bench_n${N}_c5.log 2>&1; echo "c5 rc=$?"
grep '^{' gpurun_out/bench_n${N}_c5.log | cut -c1-300; grep -i "error\|Traceback" gpurun_out/bench_n${N}_c5.log | head -5
timeout 600 bash -c "$(declare -f run); N=$N; run --workload c3 --steps 5 --warmup 3" > gpurun_out/bench_n${N}_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_n${N}_c3.log | cut -c1-300
nvidia-smi topo -m > gpurun_out/topo_n${N}.txt 2>&1

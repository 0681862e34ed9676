# round 2, run F (2 GPUs): the two-process peer-exchange test, C4 at N = 2 (headline batch + sweep), C3 at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or peer_exchange" > gpurun_out/t2_multi.log 2>&1; echo "tests rc=$?" >> gpurun_out/t2_multi.log
tail -6 gpurun_out/t2_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?" >> gpurun_out/bench_n2.log
grep -v "^\[W\|^W1" gpurun_out/bench_n2.log | tail -3 | cut -c1-5000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/bench_c3_n2.log 2>&1; echo "c3 n2 rc=$?" >> gpurun_out/bench_c3_n2.log
tail -2 gpurun_out/bench_c3_n2.log | cut -c1-1500

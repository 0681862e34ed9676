# round 2, run L (gpurun --gpus N): schedule sweep only
N=${1:-2}
GRID=${2:-0:0}
BATCHES=${3:-4096}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tune_schedule_sharded.py --batches $BATCHES --grid $GRID > gpurun_out/tune_n$N.log 2>&1; echo "tune rc=$?"
grep -v RESULT gpurun_out/tune_n$N.log | grep "^B\|rror" | cut -c1-250

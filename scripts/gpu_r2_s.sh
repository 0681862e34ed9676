# round 2, run S: extra CTAs on the 4 left-over SMs, on (0) / off (32), interleaved, 3 repetitions
mkdir -p gpurun_out
timeout 900 python scripts/ab_knobs.py --variants 0,32 --batches 4096 --reps 4 > gpurun_out/ab.log 2>&1; echo "ab rc=$?"
grep -v RESULT gpurun_out/ab.log | cut -c1-300

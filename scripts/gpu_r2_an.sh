# round 2, run AN: round-schedule sweep again (hits are cheaper with the one-chunk epilogue)
mkdir -p gpurun_out
timeout 900 python scripts/tune_schedule.py --batches 4096,64,1 --steps 20 > gpurun_out/tune.log 2>&1; echo "tune rc=$?"
grep -v "^{" gpurun_out/tune.log | tail -80

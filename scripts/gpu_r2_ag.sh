# round 2, run AG: 128-query blocks (four accumulator stages) in the hit-dense short rounds, by round length in tiles
mkdir -p gpurun_out
timeout 900 python scripts/ab_knobs.py --knob HWER_NARROW_TILES --variants 0,800,3000,10000 --batches 4096 --reps 2 > gpurun_out/ab.log 2>&1; echo "ab rc=$?"
grep -v RESULT gpurun_out/ab.log | cut -c1-300

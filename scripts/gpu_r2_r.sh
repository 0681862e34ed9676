# round 2, run R: DRAM traffic per step and launch lists (C4 only, --no-extras)
bash scripts/gpu_traffic.sh
bash scripts/gpu_list.sh

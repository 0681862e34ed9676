"""Turns the ncu csv files of scripts/gpu_traffic.sh (gpurun_out/traffic_b{4096,64,1}.csv) into the
profiles/r*_traffic.json that bench.py reads for `roofline.traffic`: DRAM bytes (read, write) per step summed over
the filter-stage launches (score_filter_tc_kernel + spill_extract_kernel).
Usage: python scripts/traffic_json.py gpurun_out profiles/r01_v16_traffic.json"""
import csv
import json
import sys

src, out = sys.argv[1], sys.argv[2]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
res = {"kernel": "score_filter_tc_kernel (+ spill_extract_kernel)",
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, bench.py --steps 2 "
              "--warmup 1, all filter-stage launches, summed per step", "batches": {}}
for B in (4096, 64, 1):
    try:
        lines = [l for l in open("%s/traffic_b%d.csv" % (src, B)) if not l.startswith("==")]
    except OSError:
        continue
    per = {}
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
        per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = v
    ids = sorted(per, key=int)
    dense = [i for i in ids if "score_filter_tc_kernel<2" in per[i]["name"] or "Li2ELi" in per[i]["name"]]
    steps = max(1, len(dense))                    # one dense (round 0) launch per step
    tot = lambda m: sum(per[i].get(m, 0.0) for i in ids)
    res["batches"][str(B)] = {
        "steps_captured": steps, "launches_per_step": len(ids) / steps,
        "dram_read_bytes_per_step": tot("dram__bytes_read.sum") / steps,
        "dram_write_bytes_per_step": tot("dram__bytes_write.sum") / steps,
        "kernel_us_per_step_under_ncu": tot("gpu__time_duration.sum") / steps,
    }
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))

"""gpurun_out/traffic_b*.csv (ncu per-launch dram bytes of the filter kernel) -> profiles/<round>_traffic.json:
bytes per STEP (sum over the step's filter launches), averaged over the steps captured."""
import csv
import json
import re
import sys


def per_step(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = {}
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        launches.setdefault(i, {"name": row["Kernel Name"]})[row["Metric Name"]] = v * mult
    seq = [launches[i] for i in sorted(launches)]
    # a step starts at each dense-round launch (template argument <2, ...>)
    starts = [i for i, l in enumerate(seq) if re.search(r"score_filter_tc_kernel<\(?int\)?2|<2,", l["name"])]
    steps = [seq[a:b] for a, b in zip(starts, starts[1:] + [len(seq)])]
    tot = lambda st, k: sum(l.get(k, 0.0) for l in st)
    n = len(steps)
    return {"steps_captured": n, "launches_per_step": len(steps[0]),
            "dram_read_bytes_per_step": sum(tot(s, "dram__bytes_read.sum") for s in steps) / n,
            "dram_write_bytes_per_step": sum(tot(s, "dram__bytes_write.sum") for s in steps) / n,
            "kernel_us_per_step_under_ncu": sum(tot(s, "gpu__time_duration.sum") for s in steps) / n}


if __name__ == "__main__":
    out = {}
    for B in (4096, 64, 1):
        try:
            out[str(B)] = per_step("gpurun_out/traffic_b%d.csv" % B)
        except Exception as e:  # noqa: BLE001
            out[str(B)] = {"error": str(e)}
    json.dump({"kernel": "score_filter_tc_kernel", "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, "
               "bench.py --steps 2 --warmup 1, all filter launches, summed per step", "batches": out},
              open(sys.argv[1] if len(sys.argv) > 1 else "profiles/traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))

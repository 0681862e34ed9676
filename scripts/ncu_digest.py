"""Text digest of an ncu report for profiles/: selected raw metrics per launch + the SASS rows with the most stall
samples.  Usage: python scripts/ncu_digest.py rep.ncu-rep [min_share_pct] > profiles/rNN_..._ncu_full_*.txt"""
import csv
import subprocess
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_src import load  # noqa: E402

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct",
        "sm__inst_issued.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled",
        "sm__warps_active.avg.pct_of_peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct")

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, body = rows[0], rows[1], rows[2:]
print("kernels:", [r[h.index("Kernel Name")][:60] for r in body])
for i, name in enumerate(h):
    if any(name.startswith(k) for k in KEEP):
        print("%-86s %-12s %s" % (name, units[i], [r[i] for r in body]))
for li, (hd, sass) in enumerate(load(rep)):
    ci = {n: i for i, n in enumerate(hd)}
    S, E, SRC = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
    tot = sum(int(r[S]) for r in sass) or 1
    print("---- launch %d: %d samples, %d SASS rows, %d warp-instructions" % (li, tot, len(sass), sum(int(r[E]) for r in sass)))
    for i, r in enumerate(sass):
        if 100.0 * int(r[S]) / tot >= min_share:
            print("  %5d %6.1f%% %11s  %s" % (i, 100.0 * int(r[S]) / tot, r[E], r[SRC][:90]))

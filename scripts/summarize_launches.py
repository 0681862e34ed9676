"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, shares, one step."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u in ("s", "second") else v
        seq.append((name, v))
    return seq


def main(path):
    seq = load(path)
    per = collections.OrderedDict()
    for n, v in seq:
        per.setdefault(n, []).append(v)
    setup = ("blend", "norm_stats", "shadow")
    tot = sum(sum(v) for k, v in per.items() if not any(s in k for s in setup))
    print("%s: %d launches; serving kernels total %.1f us" % (path, len(seq), tot))
    for k, v in per.items():
        share = "" if any(s in k for s in setup) else "share=%5.1f%%" % (sum(v) / tot * 100)
        print("  %-40s n=%4d total=%12.1f us mean=%10.1f max=%10.1f %s" % (k, len(v), sum(v), sum(v) / len(v), max(v), share))
    idx = [i for i, (n, v) in enumerate(seq) if n.startswith("fill_f32")]
    if len(idx) > 2:
        print("  one step:", [(n[:16], round(v, 1)) for n, v in seq[idx[1]:idx[2]]])


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)

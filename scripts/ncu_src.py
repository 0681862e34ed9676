"""Per-launch SASS hot spots of an ncu report: python scripts/ncu_src.py rep.ncu-rep [launch] [min_samples]"""
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    launches = []
    for k, hi in enumerate(hidx):
        end = hidx[k + 1] - 1 if k + 1 < len(hidx) else len(rows)
        launches.append((rows[hi], rows[hi + 1:end]))
    return launches[::2] if len(launches) % 2 == 0 and len(launches) > 1 else launches   # ncu prints each launch twice


if __name__ == "__main__":
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    thresh = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    h, body = load(rep)[which]
    ci = {n: i for i, n in enumerate(h)}
    S, E, SRC = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
    tot = sum(int(r[S]) for r in body)
    print("launch %d: %d samples, %d warp-instructions" % (which, tot, sum(int(r[E]) for r in body)))
    for i, r in enumerate(body):
        if int(r[S]) >= thresh:
            print("%5d %7s %11s %s" % (i, r[S], r[E], r[SRC][:100]))

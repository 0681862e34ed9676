"""Digest of an ncu --set full report: key raw metrics + stall samples by SASS region for one kernel."""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
            "lts__throughput.avg.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
            "launch__grid_size", "sm__cycles_elapsed.max", "sm__inst_issued.avg.per_cycle_active",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
            "smsp__average_warps_issue_stalled"]
    for i, h in enumerate(hdr):
        if any(w in h for w in want) and "pcsamp" not in h and "TriageCompute" not in h:
            vals = [r[i] for r in rows[2:]]
            if all(v in ("0", "") for v in vals):
                continue
            print("%-82s %-12s %s" % (h, units[i], vals))


def source(rep, top=28):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for k, hi in enumerate(hidx):
        end = hidx[k + 1] - 1 if k + 1 < len(hidx) else len(rows)
        h, body = rows[hi], rows[hi + 1:end]
        ci = {n: i for i, n in enumerate(h)}
        samp, src, ex = ci["# Samples"], ci["Source"], ci["Instructions Executed"]
        tot = sum(int(r[samp]) for r in body)
        print("---- launch %d: %d samples, %d SASS rows, %d instr executed" %
              (k, tot, len(body), sum(int(r[ex]) for r in body)))
        for a in range(0, len(body), 100):
            s = sum(int(r[samp]) for r in body[a:a + 100])
            if s > tot * 0.02:
                print("   rows %4d-%4d: %5.1f%% of samples, instr %d" % (a, a + 99, 100.0 * s / tot,
                                                                       sum(int(r[ex]) for r in body[a:a + 100])))
        order = sorted(range(len(body)), key=lambda i: -int(body[i][samp]))[:top]
        for i in sorted(order):
            print("   %5d %7s %11s %s" % (i, body[i][samp], body[i][ex], body[i][src][:90]))
        break


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1])

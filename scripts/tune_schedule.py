"""Round-schedule sweep for hwer_topk on the C4 workload (10 M x 128, top-100): times whole steps for a grid of
(first dense round rows, growth factor) through the HWER_FIRST_ROWS / HWER_GROWTH tuning knobs of
csrc/api.cu:make_schedule.  Results are checked against the default schedule's answer (same rows for every
schedule, or the schedule is wrong).  Usage: python scripts/tune_schedule.py [--batches 1,64] [--items N]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hwer_b200 as hw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batches", default="1,64")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    table = torch.empty((a.items, a.dim), dtype=torch.float32, device=dev)
    for b in range(0, a.items, 2_000_000):
        e = min(a.items, b + 2_000_000)
        table[b:e] = hw.ops.unit_length(torch.randn((e - b, a.dim), generator=g, device=dev))
    shadow = hw.ops.make_shadow(table)
    index = hw.ops.TopKIndex(table, shadow, max_norm=1.0001)
    out = {}
    for B in [int(x) for x in a.batches.split(",")]:
        q = hw.ops.unit_length(torch.randn((B, a.dim), generator=g, device=dev))
        for v in ("HWER_FIRST_ROWS", "HWER_GROWTH"):
            os.environ.pop(v, None)
        ref_idx = index.topk(q, a.k)[0].clone()
        grid = [(None, None)] + [(f, gr) for f in (2048, 4096, 8192, 16384) for gr in ((2, 3, 4, 6, 8) if B > 512 else (2, 4, 8, 16, 32, 64))]
        for first, growth in grid:
            for v, val in (("HWER_FIRST_ROWS", first), ("HWER_GROWTH", growth)):
                if val is None:
                    os.environ.pop(v, None)
                else:
                    os.environ[v] = str(val)
            index = hw.ops.TopKIndex(table, shadow, max_norm=1.0001)   # the knobs are read when an index is created
            try:
                for _ in range(3):
                    idx = index.topk(q, a.k)[0]
                assert torch.equal(idx, ref_idx), "schedule changed the answer"
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.steps):
                    index.topk_async(q, a.k)
                e1.record()
                torch.cuda.synchronize()
                rc, need = index.finish()
                ms = e0.elapsed_time(e1) / a.steps
                out["B%d first=%s g=%s" % (B, first, growth)] = ms if rc == 0 else "overflow(%d)" % need
            except Exception as ex:      # a knob combination the library refuses
                out["B%d first=%s g=%s" % (B, first, growth)] = "error: %s" % str(ex)[:80]
            print(B, first, growth, out["B%d first=%s g=%s" % (B, first, growth)], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

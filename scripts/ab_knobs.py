"""Same-process A/B of kernel features on the C4 workload (10 M x 128, top-100) through the HWER_DISABLE knob
(csrc/api.cu: bit 0 = filter with the bias MMA instead of scaled queries, bit 1 = old single-shape final kernel).
Variants are interleaved (A B A B) inside one process: different boxes differ by a few per cent, and a
power-capped GPU drifts during a run.  Usage: python scripts/ab_knobs.py [--variants 3,2,1,0] [--batches 4096,64,1]"""
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
    ap.add_argument("--variants", default="3,0")
    ap.add_argument("--batches", default="4096,64,1")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--knob", default="HWER_DISABLE", help="the environment knob the variants are values of")
    ap.add_argument("--env", default="", help="extra knob per variant, e.g. 'HWER_GROWTH=4' applied to all")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    n, d, k = 10_000_000, 128, 100
    table = torch.empty((n, d), dtype=torch.float32, device=dev)
    for b in range(0, n, 2_000_000):
        table[b:b + 2_000_000] = hw.ops.unit_length(torch.randn((2_000_000, d), generator=g, device=dev))
    shadow = hw.ops.make_shadow(table)
    for kv in [x for x in a.env.split(",") if x]:
        key, val = kv.split("=")
        os.environ[key] = val
    res = {}
    ref = {}
    for rep in range(a.reps):
        for v in a.variants.split(","):
            os.environ[a.knob] = v
            index = hw.ops.TopKIndex(table, shadow, max_norm=1.0001)
            for B in [int(x) for x in a.batches.split(",")]:
                steps = 8 if B >= 1024 else 40
                gq = torch.Generator(device=dev).manual_seed(100 + B)
                q = hw.ops.unit_length(torch.randn((B, d), generator=gq, device=dev))
                for _ in range(3):
                    idx = index.topk(q, k)[0]
                chk = int(idx.sum().item())
                assert ref.setdefault(B, chk) == chk, "variant %s changed the answer at B=%d" % (v, B)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(steps):
                    index.topk_async(q, k)
                e1.record()
                torch.cuda.synchronize()
                index.finish()
                step_ms = e0.elapsed_time(e1) / steps
                index.profile(True)
                for _ in range(steps):
                    index.topk_async(q, k)
                torch.cuda.synchronize()
                per = index.profile_launches()
                stages = index.profile_stages()
                index.profile_read()
                index.profile(False)
                r = len(per) // steps
                rounds = [sum(per[s * r + i] for s in range(steps)) / steps * 1e3 for i in range(r)]
                rec = {"step_ms": round(step_ms, 4), "rounds_us": [round(x, 1) for x in rounds],
                       "stages_ms": {kk: round(vv / steps, 4) for kk, vv in stages.items()}}
                res.setdefault("v%s B%d" % (v, B), []).append(rec)
                print("v%s B%d rep%d %s" % (v, B, rep, json.dumps(rec)), flush=True)
            del index
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()

"""Same-box A/B of library builds: per-round score-filter times and whole-step time on the C4 workload.
Different gpurun boxes differ by a few per cent (clocks, power cap), so variants are only comparable inside one run;
each variant runs in its own process, interleaved (A B A B).
Usage: python scripts/ab_rounds.py variants/libhwer_b200_a.so variants/libhwer_b200_b.so [...]   (worker: --worker)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    sys.path.insert(0, ROOT)
    import torch
    import hwer_b200 as hw
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    n, d, k = 10_000_000, 128, 100
    table = torch.empty((n, d), dtype=torch.float32, device=dev)
    for b in range(0, n, 2_000_000):
        table[b:b + 2_000_000] = hw.ops.unit_length(torch.randn((2_000_000, d), generator=g, device=dev))
    index = hw.ops.TopKIndex(table, hw.ops.make_shadow(table), max_norm=1.0001)
    out = {}
    for B, steps in ((4096, 8), (64, 40), (1, 40)):
        q = hw.ops.unit_length(torch.randn((B, d), generator=g, device=dev))
        for _ in range(3):
            idx = index.topk(q, k)[0]
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
        st = {k_: round(v_ / steps, 4) for k_, v_ in index.profile_stages().items()}
        index.profile_read()
        index.profile(False)
        r = len(per) // steps
        rounds = [sum(per[s * r + i] for s in range(steps)) / steps * 1e3 for i in range(r)]
        out[str(B)] = {"step_ms": step_ms, "rounds_us": [round(x, 1) for x in rounds], "filter_ms": sum(rounds) / 1e3,
                       "checksum": int(idx.sum().item()), "stages": st}
    print("RESULT " + json.dumps(out), flush=True)


def main():
    libs = sys.argv[1:]
    res = {l: [] for l in libs}
    for rep in range(2):
        for l in libs:
            env = dict(os.environ, HWER_B200_LIB=os.path.join(ROOT, l))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True, text=True)
            line = [x for x in r.stdout.splitlines() if x.startswith("RESULT ")]
            if not line:
                print(l, "FAILED", r.stderr[-800:])
                continue
            res[l].append(json.loads(line[0][7:]))
    for B in ("4096", "64", "1"):
        for l in libs:
            for x in res[l]:
                print("B=%-4s %-34s step %.3f ms filter %.3f ms rounds %s chk %d stages %s" % (
                    B, os.path.basename(l), x[B]["step_ms"], x[B]["filter_ms"], x[B]["rounds_us"], x[B]["checksum"],
                    x[B].get("stages")))


if __name__ == "__main__":
    worker() if "--worker" in sys.argv else main()

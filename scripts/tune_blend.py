"""blend_normalize sweep on the C4 shape (10 M x 128): rows per warp and CTAs per SM through the HWER_BLEND_RPW /
HWER_BLEND_CTAS knobs.  The knobs are read once per process, so every cell runs in its own process.
Usage: python scripts/tune_blend.py   (worker: --worker)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    sys.path.insert(0, ROOT)
    import torch
    import hwer_b200 as hw
    n, d = 10_000_000, 128
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.randn((n, d), generator=g, device=dev)
    x = torch.randn((n, d), generator=g, device=dev)
    for _ in range(2):
        t, s = hw.ops.blend_normalize(c, x, 0.5)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        del t, s
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t, s = hw.ops.blend_normalize(c, x, 0.5)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    byt = n * d * 12 + n * s.shape[1] * 2
    a = torch.empty(1 << 29, dtype=torch.float32, device=dev)
    b = torch.empty_like(a)
    cb = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        cb = min(cb, e0.elapsed_time(e1))
    print("RESULT " + json.dumps({"ms": best, "gbs": byt / best / 1e6, "copy_gbs": 2 * a.numel() * 4 / cb / 1e6,
                                  "checksum": float(t[::100003].double().sum().item())}))


def main():
    for rpw in (0, 1, 4):
        for ctas in (8, 16, 32):
            env = dict(os.environ, HWER_BLEND_RPW=str(rpw), HWER_BLEND_CTAS=str(ctas))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True, text=True)
            line = [x for x in r.stdout.splitlines() if x.startswith("RESULT ")]
            print("rpw=%d ctas=%d %s" % (rpw, ctas, line[0][7:] if line else "FAILED " + r.stderr[-300:]), flush=True)


if __name__ == "__main__":
    worker() if "--worker" in sys.argv else main()

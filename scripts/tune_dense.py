"""Sweep of the converged-hit-path switch (HWER_DENSE_LANES) on the C4 workload: whole-step and filter-kernel time
per setting, B = 4096 and 64.  Usage: python scripts/tune_dense.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hwer_b200 as hw  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
n, d, k = 10_000_000, 128, 100
table = torch.empty((n, d), dtype=torch.float32, device=dev)
for b in range(0, n, 2_000_000):
    table[b:b + 2_000_000] = hw.ops.unit_length(torch.randn((2_000_000, d), generator=g, device=dev))
index = hw.ops.TopKIndex(table, hw.ops.make_shadow(table), max_norm=1.0001)
for B in (4096, 64):
    q = hw.ops.unit_length(torch.randn((B, d), generator=g, device=dev))
    os.environ.pop("HWER_DENSE_LANES", None)
    ref = index.topk(q, k)[0].clone()
    for rep in range(2):
        for v in (33, 4, 6, 8, 12, 16, 24):
            os.environ["HWER_DENSE_LANES"] = str(v)
            for _ in range(2):
                idx = index.topk(q, k)[0]
            assert torch.equal(idx, ref)
            index.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(8):
                index.topk_async(q, k)
            e1.record()
            torch.cuda.synchronize()
            index.finish()
            filt_ms, fl, ol = index.profile_read()
            index.profile(False)
            print("B=%d dense_lanes=%d step %.3f ms filter %.3f ms" % (B, v, e0.elapsed_time(e1) / 8, filt_ms / 8), flush=True)

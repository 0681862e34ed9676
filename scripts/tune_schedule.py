"""GPU tuning aid: time hwer_topk on the C4 table for several round schedules (HWER_GROWTH / HWER_LATE_ROWS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hwer_b200 as hw  # noqa: E402

n, d, k = int(os.environ.get("N", 10_000_000)), 128, 100
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
table, shadow = hw.ops.blend_normalize(torch.randn((n, d), generator=g, device=dev),
                                       torch.randn((n, d), generator=g, device=dev), 0.5)
index = hw.ops.TopKIndex(table, shadow, max_norm=hw.ops.norm_stats(table)[4])


def run(B, growth, late, steps):
    os.environ["HWER_GROWTH"] = str(growth)
    os.environ["HWER_LATE_ROWS"] = str(late)
    q = hw.ops.unit_length(torch.randn((B, d), generator=g, device=dev))
    for _ in range(3):
        index.topk_async(q, k)
    rc, need = index.finish()
    assert rc == 0, (rc, need)
    index.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        index.topk_async(q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fms, fl, ol = index.profile_read()
    index.profile(False)
    print("B=%5d growth=%2d late=%9d : %8.3f ms/step  %10.0f q/s  filter %7.3f ms  launches/step %.0f" %
          (B, growth, late, ms, B / ms * 1e3, fms / steps, (fl + ol) / steps), flush=True)


for B, steps in ((1, 50), (16, 50), (64, 50), (256, 30), (1024, 20), (4096, 10)):
    for growth, late in ((16, 1 << 40), (16, 262144), (16, 1048576), (8, 262144), (8, 1048576), (4, 1 << 40), (32, 1 << 40), (32, 1048576)):
        if 3 * k * growth > 16384:
            continue
        run(B, growth, late, steps)

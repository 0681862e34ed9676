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


def run(B, growth, late, steps, first=1024):
    os.environ["HWER_FIRST_ROWS"] = str(first)
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
    print("B=%5d first=%5d growth=%2d late=%9d : %8.3f ms/step  %10.0f q/s  filter %7.3f ms  launches/step %.0f" %
          (B, first, growth, late, ms, B / ms * 1e3, fms / steps, (fl + ol) / steps), flush=True)


CONFIGS = {
    4096: [(1024, 8, 262144), (8192, 8, 262144), (8192, 4, 262144), (8192, 2, 1 << 40), (8192, 1, 0), (4096, 1, 0),
           (16384, 1, 0), (8192, 4, 65536), (4096, 2, 1 << 40), (2048, 2, 1 << 40)],
    1024: [(1024, 8, 262144), (8192, 8, 262144), (8192, 2, 1 << 40), (8192, 1, 0), (8192, 4, 65536)],
    256: [(1024, 8, 262144), (8192, 8, 262144), (8192, 4, 1 << 40), (8192, 2, 1 << 40), (8192, 8, 1 << 40)],
    64: [(1024, 8, 1048576), (8192, 8, 1048576), (8192, 8, 1 << 40), (8192, 16, 1 << 40), (8192, 32, 1 << 40), (16384, 32, 1 << 40)],
    16: [(1024, 32, 1 << 40), (8192, 32, 1 << 40), (16384, 32, 1 << 40)],
    1: [(1024, 32, 1 << 40), (8192, 32, 1 << 40), (16384, 32, 1 << 40)],
}
for B, steps in ((4096, 8), (1024, 15), (256, 20), (64, 40), (16, 40), (1, 40)):
    for first, growth, late in CONFIGS[B]:
        if 3 * k * growth > 16384:
            continue
        run(B, growth, late, steps, first)

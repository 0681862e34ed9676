"""NCF re-rank throughput: B * k (anchor, candidate) pairs through hwer_ncf_score, CUDA events on the
launching stream, next to the same MLP in torch fp32 on the host cores (what the reference runs, gcn_ncf.py:344-359).
Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hwer_b200 as hw  # noqa: E402
import hwer_oracle as O  # noqa: E402

B, k, F, depth, n = 4096, 200, 128, 3, 1_000_000
P = B * k
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn((n + 1, F), generator=g, device=dev) * 0.3
dims = O.ncf_layer_dims(F, depth)
params = torch.cat([torch.cat([(torch.randn((o, i), generator=g, device=dev) / i ** 0.5).reshape(-1),
                               torch.randn((o,), generator=g, device=dev) * 0.1]) for i, o in dims])
src = torch.randint(1, n + 1, (B,), generator=g, device=dev).repeat_interleave(k)
dst = torch.randint(1, n + 1, (P,), generator=g, device=dev)
for _ in range(2):
    out = hw.ops.ncf_score(h, params, src, dst, depth)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
e0.record()
for _ in range(steps):
    out = hw.ops.ncf_score(h, params, src, dst, depth)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flops = 2.0 * P * sum(i * o for i, o in dims)
# host: the reference's batches of 512 pairs (gcn_ncf.py:338,344-359) on a bounded sample
sample = 20480
hc, pc = h.cpu(), params.cpu()
ws, off = [], 0
for i, o in dims:
    ws.append((pc[off:off + i * o].reshape(o, i), pc[off + i * o:off + i * o + o]))
    off += i * o + o
sc, dc = src[:sample].cpu(), dst[:sample].cpu()
t0 = time.time()
with torch.no_grad():
    res = []
    for a, b in zip(sc.split(512), dc.split(512)):
        x = torch.cat([hc[a], hc[b]], 1)
        for li, (w, bias) in enumerate(ws):
            x = torch.nn.functional.linear(x, w, bias)
            x = torch.nn.functional.leaky_relu(x, 0.01) if li < len(ws) - 1 else torch.sigmoid(x)
        res.append(x.flatten())
cpu_s = time.time() - t0
err = float((torch.cat(res) - out[:sample].cpu()).abs().max())
print(json.dumps({"op": "ncf_rerank", "pairs": P, "F": F, "depth": depth, "ms": ms, "pairs_per_s": P / ms * 1e3,
                  "tflops": flops / ms / 1e9, "ffma_peak_tflops_at_1965MHz": 148 * 128 * 2 * 1.965e9 / 1e12,
                  "cpu_pairs_per_s": sample / cpu_s, "cpu_threads": torch.get_num_threads(),
                  "max_abs_diff_vs_torch_cpu": err}))

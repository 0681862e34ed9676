"""Debug aid: world ranks in one process; compare the exchange result with merge_topk over plain local searches."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hwer_b200 as hw  # noqa: E402
from hwer_b200 import _native as N  # noqa: E402

lib = N.lib()
world, B, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sync_phases = len(sys.argv) > 4 and sys.argv[4] == "sync"
n, d = 30000, 128
rs = np.random.RandomState(91)
t_np = rs.standard_normal((n, d)).astype(np.float32)
t_np /= np.linalg.norm(t_np, axis=1, keepdims=True)
t = torch.from_numpy(t_np).cuda()
q = torch.from_numpy(rs.standard_normal((B, d)).astype(np.float32)).cuda()
q = q / q.norm(dim=1, keepdim=True)
whole = hw.ops.TopKIndex(t)
ref_idx, ref_sc, ref_s64 = whole.topk(q, k, want_f64=True)


class Ex:
    def __init__(self, h):
        self._h = h


nbytes = lib.hwer_exchange_bytes(world, 256, k)
bases = (ctypes.c_void_p * world)()
handle = (ctypes.c_ubyte * 64)()
for r in range(world):
    p = ctypes.c_void_p()
    N.check(lib.hwer_peer_alloc(nbytes, ctypes.byref(p), handle))
    bases[r] = p
ex, shards, streams, local = [], [], [], []
for r in range(world):
    h = ctypes.c_void_p()
    N.check(lib.hwer_exchange_create(ctypes.byref(h), world, r, 256, k, bases, 0))
    ex.append(Ex(h))
    b, e = hw.sharded.partition(n, world, r)
    ix = hw.ops.TopKIndex(t[b:e].contiguous())
    li, ls, ls64 = ix.topk(q, k, idx_offset=b, want_f64=True)
    local.append((li, ls64))
    shards.append((ix, b))
    streams.append(torch.cuda.Stream())
mi, ms, ms64 = hw.ops.merge_topk(torch.stack([l[1] for l in local]).contiguous(), torch.stack([l[0] for l in local]).contiguous(), want_f64=True)
print("merge_topk over plain local searches == whole:", torch.equal(mi, ref_idx))
outs = [(torch.empty((B, k), dtype=torch.int64, device="cuda"), torch.empty((B, k), dtype=torch.float32, device="cuda"),
         torch.empty((B, k), dtype=torch.float64, device="cuda")) for _ in range(world)]
torch.cuda.synchronize()
for phase in (1, 2, 4):
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            shards[r][0].topk_sharded_async(ex[r], q, k, idx_offset=shards[r][1], phases=phase, out=outs[r])
        if len(sys.argv) > 5 and sys.argv[5] == "serial":
            torch.cuda.synchronize()
    if sync_phases:
        torch.cuda.synchronize()
torch.cuda.synchronize()
for r in range(world):
    print("rank", r, "err", lib.hwer_exchange_error(ex[r]._h, None), "equal", torch.equal(outs[r][0], ref_idx))
bad = (outs[0][0] != ref_idx).any(dim=1).nonzero().flatten().tolist()
for qq in bad[:3]:
    print("query", qq, "got", outs[0][0][qq].tolist(), "\n   want", ref_idx[qq].tolist())
    for r in range(world):
        print("   shard", r, "local", local[r][0][qq].tolist())
# raw exchange buffer of rank 0: xs/xi as the merge saw them
L_flags = 256
q_cap = (256 + world - 1) // world
xbytes = (world * q_cap * k * 8 + 255) // 256 * 256
raw = (ctypes.c_ubyte * nbytes)()
import ctypes as C
cudart = C.CDLL("libcudart.so")
cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
print("memcpy rc", cudart.cudaMemcpy(raw, bases[0], nbytes, 2))
buf = np.frombuffer(raw, dtype=np.uint8)
xs = buf[L_flags:L_flags + world * q_cap * k * 8].view(np.float64).reshape(world, q_cap, k)
xi = buf[L_flags + xbytes:L_flags + xbytes + world * q_cap * k * 8].view(np.int64).reshape(world, q_cap, k)
print("flags", buf[:128].view(np.uint32)[:20])
for qq in bad[:2]:
    if qq < (B + world - 1) // world:
        for g in range(world):
            print("   q", qq, "xbuf src", g, xi[g, qq][:12].tolist(), np.round(xs[g, qq][:12], 4).tolist())
            print("        local     ", local[g][0][qq][:12].tolist(), np.round(local[g][1][qq][:12].cpu().numpy(), 4).tolist())

import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hwer_b200 as hw
n, d, B, k = 10000, 128, int(sys.argv[1]), int(sys.argv[2])
rs = np.random.RandomState(91)
t_np = rs.standard_normal((30000, d)).astype(np.float32)
t_np /= np.linalg.norm(t_np, axis=1, keepdims=True)
t = torch.from_numpy(t_np).cuda()
q = torch.from_numpy(rs.standard_normal((B, d)).astype(np.float32)).cuda()
q = q / q.norm(dim=1, keepdim=True)
for shard in range(3):
    ix = hw.ops.TopKIndex(t[shard * n:(shard + 1) * n].contiguous())
    ref = ix.topk(q, k)[0]
    for trial in range(3):
        a = ix.topk(q, k)[0]
        print("shard", shard, "default stream again equal:", torch.equal(a, ref))
    s = torch.cuda.Stream()
    for trial in range(3):
        with torch.cuda.stream(s):
            a, _, _ = ix.topk_async(q, k)
            rc, need = ix.finish()
        torch.cuda.synchronize()
        print("shard", shard, "side stream equal:", torch.equal(a, ref), rc, need)

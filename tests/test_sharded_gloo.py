"""CPU, world_size 2, gloo: the multi-GPU exchange step (pad -> one all-gather -> merge order) checked against
the oracle's single-shard answer.  The merge arithmetic itself is a CUDA kernel (covered by -m gpu tests); here a
numpy lexsort stands in as the checker so the collective plumbing is exercised without a GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, d, B, k, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hwer_oracle as O
    from hwer_b200.sharded import gather_shard_results, pad_local_result, partition
    rs = np.random.RandomState(5)
    table = O.unit_length(rs.standard_normal((n, d)).astype(np.float32), axis=1)
    table[n - 3:] = table[:3]                       # duplicates across the shard boundary exercise the tie rule
    q = O.unit_length(rs.standard_normal((B, d)).astype(np.float32), axis=1)
    b, e = partition(n, world, rank)
    kl = min(k, e - b)
    idx, sc = O.exact_topk(table[b:e], q, kl)       # the oracle plays the per-GPU search
    idx_t, s64_t = pad_local_result(torch.from_numpy(idx + b), torch.from_numpy(sc), k)
    gs, gi = gather_shard_results(idx_t, s64_t)
    assert gs.shape == (world, B, k) and gi.shape == (world, B, k) and gs.dtype == torch.float64
    assert torch.equal(gi[rank], idx_t) and torch.equal(gs[rank], s64_t)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), gs=gs.numpy(), gi=gi.numpy(), table=table, q=q)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_merge_order(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwer_oracle as O
    n, d, B, k = 1000, 16, 9, 20
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n, d, B, k, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    gs, gi = z["gs"], z["gi"]
    ref_idx, ref_sc = O.exact_topk(z["table"], z["q"], k)
    for r in range(B):
        s = gs[:, r].reshape(-1)
        i = gi[:, r].reshape(-1)
        keep = i >= 0
        order = np.lexsort((i[keep], -s[keep]))[:k]          # (score desc, row asc): the merge kernel's rule
        np.testing.assert_array_equal(i[keep][order], ref_idx[r])
        np.testing.assert_allclose(s[keep][order], ref_sc[r], atol=1e-12)


def test_small_shard_is_padded():
    from hwer_b200.sharded import pad_local_result
    idx = torch.arange(6).reshape(2, 3)
    s = torch.rand(2, 3, dtype=torch.float64)
    pi, ps = pad_local_result(idx, s, 5)
    assert pi.shape == (2, 5) and torch.all(pi[:, 3:] == -1) and torch.all(torch.isinf(ps[:, 3:]))


def _query_worker(rank, world, port, n, d, B, k, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hwer_oracle as O
    from hwer_b200.sharded import gather_query_shards, partition
    rs = np.random.RandomState(6)
    table = O.unit_length(rs.standard_normal((n, d)).astype(np.float32), axis=1)
    q = O.unit_length(rs.standard_normal((B, d)).astype(np.float32), axis=1)
    b, e = partition(B, world, rank)
    idx, sc = O.exact_topk(table, q[b:e], k)        # the oracle plays this rank's search over the replicated table
    gi, gs = gather_query_shards(torch.from_numpy(idx), torch.from_numpy(sc.astype(np.float32)), B)
    assert gi.shape == (B, k) and gs.shape == (B, k) and gs.dtype == torch.float32
    ref_idx, ref_sc = O.exact_topk(table, q, k)
    np.testing.assert_array_equal(gi.numpy(), ref_idx)
    np.testing.assert_array_equal(gs.numpy(), ref_sc.astype(np.float32))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_query_sharding_gathers_in_query_order(tmp_path):
    n, d, B, k = 500, 16, 11, 7                     # 11 queries over 2 ranks: unequal slices (5 and 6)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_query_worker, args=(2, port, n, d, B, k, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))

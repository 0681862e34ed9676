"""Item-sharded top-k across the GPUs of one node (no reference counterpart -- the
reference is single-process).

GPU g owns the contiguous catalogue rows [g*N/G, (g+1)*N/G); queries are replicated; every rank computes its
local exact top-k (rows offset to global ids, fp64 scores), one all-gather exchanges the [B, k] results, and the
merge kernel applies the same (score desc, row asc) rule -- so 1, 2, 4 and 8-GPU runs return identical results.

Small catalogues served to ALL users (configs C2 / C3: the item table is <= 28 MB) shard the other way:
`QueryShardedTopK` replicates the table, gives every rank a contiguous slice of the queries and needs no
collective on the data path -- only an optional all-gather of the finished [B, k] results.
"""
import ctypes
from typing import Tuple

import torch
import torch.distributed as dist

from . import _native as N
from . import ops


def partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range owned by `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def pad_local_result(idx, s64, k):
    """Local shards smaller than k return fewer columns; pad with (row -1, score -inf)."""
    B, kl = idx.shape
    if kl == k:
        return idx, s64
    pi = torch.full((B, k), -1, dtype=idx.dtype, device=idx.device)
    ps = torch.full((B, k), float("-inf"), dtype=s64.dtype, device=s64.device)
    pi[:, :kl] = idx
    ps[:, :kl] = s64
    return pi, ps


def gather_shard_results(idx, s64, group=None):
    """One collective for both arrays: [B, k] per rank -> ([G, B, k] fp64 scores, [G, B, k] int64 rows)."""
    world = dist.get_world_size(group)
    packed = torch.stack([s64.view(torch.int64), idx], dim=0).contiguous()      # [2, B, k] int64
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    allp = torch.stack(out, dim=0)                                               # [G, 2, B, k]
    return allp[:, 0].contiguous().view(torch.float64), allp[:, 1].contiguous()


class PeerExchange:
    """The exchange buffers of hwer_topk_sharded (csrc/exchange.cu): one cudaMalloc'd buffer per rank, opened on
    every other rank through CUDA IPC, so kernels store results straight into the peer that merges them.
    torch.distributed only carries the 64-byte handles (host side, once)."""

    def __init__(self, b_cap: int, k_cap: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.b_cap, self.k_cap = int(b_cap), int(k_cap)
        self.device = torch.device(device)
        lib = N.lib()
        nbytes = lib.hwer_exchange_bytes(self.world, self.b_cap, self.k_cap)
        if nbytes <= 0:
            raise ValueError("bad exchange shape (world %d, batch %d, k %d)" % (self.world, b_cap, k_cap))
        self._own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * N.IPC_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            N.check(lib.hwer_peer_alloc(nbytes, ctypes.byref(self._own), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._opened = []
            bases = (ctypes.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    bases[r] = self._own
                    continue
                ptr = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * N.IPC_HANDLE_BYTES).from_buffer_copy(h)
                N.check(lib.hwer_peer_open(buf, ctypes.byref(ptr)))
                self._opened.append(ptr)
                bases[r] = ptr
            self._h = ctypes.c_void_p()
            N.check(lib.hwer_exchange_create(ctypes.byref(self._h), self.world, self.rank, self.b_cap, self.k_cap, bases,
                                             self.device.index or 0))
        dist.barrier(group=group)          # every rank has opened every buffer before anyone stores into one

    def owner_range(self, n_queries: int) -> Tuple[int, int]:
        """[lo, hi) of the queries of a batch of `n_queries` that this rank merges (csrc/exchange.cu: contiguous
        ranges of ceil(B / world) queries)."""
        per = (int(n_queries) + self.world - 1) // self.world
        lo = min(self.rank * per, int(n_queries))
        return lo, min(lo + per, int(n_queries))

    def configure(self, largest_shard_rows: int, share_thresholds: bool = True):
        N.check(N.lib().hwer_exchange_configure(self._h, int(largest_shard_rows), 1 if share_thresholds else 0))

    def check(self):
        """Synchronises the current stream; raises if a wait on a peer GPU timed out."""
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_exchange_error(self._h, ops._stream(self.device)))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if not h:
            return
        lib = N.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            lib.hwer_exchange_destroy(h)
            for p in self._opened:
                lib.hwer_peer_close(p)
        try:
            dist.barrier(group=self.group)  # nobody frees a buffer a peer still has open
        except Exception:
            pass
        with torch.cuda.device(self.device):
            lib.hwer_peer_free(self._own)

    def __del__(self):
        # freeing needs a collective; without an explicit close() the buffers live until the process exits
        pass


class ShardedTopK:
    """One rank's share of an item-sharded catalogue.  `topk` returns the merged global result on every rank:
    through the peer-memory exchange (`exchange="p2p"`, the default on a multi-GPU node) or through one NCCL
    all-gather + local merge (`exchange="nccl"`)."""

    def __init__(self, local_table, row_offset: int, shadow=None, group=None, max_norm=None, exchange="auto",
                 share_thresholds=True):
        self.group = group
        self.row_offset = int(row_offset)
        self.rows_max = int(local_table.shape[0])
        self.rows_min = self.rows_max        # smallest shard over all ranks: every rank must pick the same exchange path
        if dist.is_initialized() and dist.get_world_size(group) > 1 and local_table.is_cuda:
            # the admission margin scales with the largest row norm of the WHOLE catalogue, and all ranks walk the
            # round schedule of the largest shard (hwer_exchange_configure): agree on both once
            if max_norm is None:
                max_norm = ops.norm_stats(local_table)[4]
            t = torch.tensor([float(max_norm), float(self.rows_max), -float(self.rows_max)], dtype=torch.float64,
                             device=local_table.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            max_norm, self.rows_max, self.rows_min = float(t[0].item()), int(t[1].item()), int(-t[2].item())
        self.index = ops.TopKIndex(local_table, shadow, max_norm=max_norm)
        self.exchange = exchange
        self.share_thresholds = bool(share_thresholds)
        self._px = None

    def local_topk(self, queries, k, mode="exact"):
        kl = min(int(k), self.index.n)
        idx, _, s64 = self.index.topk(queries, kl, mode, idx_offset=self.row_offset, want_f64=True)
        return pad_local_result(idx, s64, int(k))

    def _peer_exchange(self, B, k):
        if self._px is None or self._px.b_cap < B or self._px.k_cap < k:
            if self._px is not None:
                self._px.close()
            self._px = PeerExchange(max(B, 256), k, self.index.device, self.group)
            self._px.configure(self.rows_max, self.share_thresholds)
        return self._px

    def topk_p2p_async(self, queries, k, mode="exact", cap=0, want_f64=False, owned=False):
        """Enqueues the fused search + peer exchange; results are valid after `index.finish()`.  owned=True: only
        the rows of the queries this rank merged (`owner_range`), see HWER_PHASE_OWNED."""
        B, k = queries.shape[0], int(k)
        px = self._peer_exchange(B, k)
        return self.index.topk_sharded_async(px, queries, k, mode, idx_offset=self.row_offset, cap=cap,
                                             want_f64=want_f64,
                                             phases=N.PHASE_ALL | (N.PHASE_OWNED if owned else 0))

    def owner_range(self, n_queries: int) -> Tuple[int, int]:
        """Queries [lo, hi) whose merged result `topk(..., owned=True)` returns on this rank."""
        n_queries = int(n_queries)
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return 0, n_queries
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        per = (n_queries + world - 1) // world
        lo = min(rank * per, n_queries)
        return lo, min(lo + per, n_queries)

    def topk(self, queries, k, mode="exact", owned=False):
        """Global top-k of all shards.  Default: the full [B, k] result on every rank.  owned=True: every rank gets
        the rows of `owner_range(B)` only -- each query's answer lands on exactly one rank (the form a serving tier
        wants: 1/world of the result to copy to each host, nothing delivered twice over NVLink)."""
        k = int(k)
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi and self.exchange in ("auto", "p2p") and self.rows_min >= k and queries.is_cuda:
            cap = 0
            for _ in range(6):
                idx, score, _ = self.topk_p2p_async(queries, k, mode, cap=cap, owned=owned)
                # one stream synchronisation per step: the largest capacity demand of ANY rank and a peer timeout
                # come back through the exchange buffers (csrc/exchange.cu), so every rank takes the same decision
                # here without a host-side collective
                rc, need = self.index.finish()
                if rc == N.HWER_OK:
                    return idx, score
                if rc != N.HWER_E_OVERFLOW:
                    N.check(rc)
                if need > ops.TopKIndex.MAX_CAP:
                    raise N.HwerError(N.HWER_E_OVERFLOW, "a query needs a candidate list of %d rows (more than %d rows "
                                      "inside the bf16 margin of its k-th score): answer it with "
                                      "TopKIndex.topk_exhaustive on every shard and merge" % (need, ops.TopKIndex.MAX_CAP))
                cap = need
            raise N.HwerError(N.HWER_E_OVERFLOW, "candidate lists kept overflowing")
        idx, s64 = self.local_topk(queries, k, mode)
        if not multi:
            return idx, s64.float()
        gs, gi = gather_shard_results(idx, s64, self.group)
        idx, score = ops.merge_topk(gs, gi)
        if owned:
            lo, hi = self.owner_range(queries.shape[0])
            return idx[lo:hi], score[lo:hi]
        return idx, score

    def close(self):
        if self._px is not None:
            self._px.close()
            self._px = None


# --------------------------------------------------------------------------- query sharding (replicated catalogue)
def gather_query_shards(idx, score, n_queries: int, group=None):
    """All-gather of per-rank result slices ([B_r, k] rows + scores, B_r = partition(n_queries, G, r)) into the
    full [n_queries, k] arrays, in query order, on every rank.  One collective: slices are padded to the largest."""
    world = dist.get_world_size(group)
    k = idx.shape[1]
    rows_max = max(partition(n_queries, world, r)[1] - partition(n_queries, world, r)[0] for r in range(world))
    packed = torch.zeros((2, rows_max, k), dtype=torch.int64, device=idx.device)
    packed[0, :idx.shape[0]] = idx
    packed[1, :idx.shape[0]] = score.double().view(torch.int64)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    idx_parts, sc_parts = [], []
    for r in range(world):
        b, e = partition(n_queries, world, r)
        idx_parts.append(out[r][0, :e - b])
        sc_parts.append(out[r][1, :e - b].contiguous().view(torch.float64))
    return torch.cat(idx_parts, dim=0), torch.cat(sc_parts, dim=0).to(score.dtype)


class QueryShardedTopK:
    """All-users retrieval over a catalogue that every GPU holds in full: rank r answers queries
    [r*B/G, (r+1)*B/G).  Results are bit-identical to one GPU answering all B (each query's search is independent
    of the others)."""

    def __init__(self, table, shadow=None, group=None, max_norm=None):
        self.group = group
        self.index = ops.TopKIndex(table, shadow, max_norm=max_norm)

    def local_slice(self, n_queries: int) -> Tuple[int, int]:
        if not dist.is_initialized():
            return 0, n_queries
        return partition(n_queries, dist.get_world_size(self.group), dist.get_rank(self.group))

    def topk(self, queries, k, mode="exact", gather=True):
        """queries: the FULL [B, d] batch (replicated, e.g. all user rows).  Returns this rank's slice when
        gather=False, else the full [B, k] result on every rank."""
        B = queries.shape[0]
        b, e = self.local_slice(B)
        idx, score = self.index.topk(queries[b:e].contiguous(), int(k), mode)[:2]
        if not gather or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return idx, score
        return gather_query_shards(idx, score, B, self.group)

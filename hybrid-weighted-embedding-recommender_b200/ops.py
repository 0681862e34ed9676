"""Tensor-level operators of the serving hot path.

Each function hands raw device pointers of torch CUDA tensors to the C ABI (include/hwer_b200.h) on torch's
current stream.  torch supplies memory and streams only; every computation below runs in libhwer_b200.so.
There is deliberately no CPU implementation: CPU tensors are rejected.
"""
import ctypes
from ctypes import c_uint32, c_void_p

import numpy as np
import torch

from . import _native as N


def _dev_ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need(t, dtype, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: hwer_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must be %d-dimensional" % (name, ndim))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous (row-major)" % name)
    return t


def shadow_width(d):
    return (int(d) + 63) // 64 * 64


def blend_normalize(content, collab, alpha=0.5, want_shadow=True):
    """V = unit(alpha * unit(content) + (1 - alpha) * unit(collab)) row-wise (the blend slot of
    GcnNCF.prepare_for_knn, hwer/gcn_ncf.py:447-456; alpha = 0 or content=None is the reference's behaviour).
    `alpha` is a float or an [N] fp32 CUDA tensor (per-row alpha).  Returns (fp32 table, bf16 shadow | None)."""
    collab = _need(collab, torch.float32, "collab", 2)
    n, d = collab.shape
    if content is not None:
        content = _need(content, torch.float32, "content", 2)
        if tuple(content.shape) != (n, d):
            raise ValueError("content %s and collaborative %s tables must have the same shape"
                             % (tuple(content.shape), (n, d)))
    alpha_rows = None
    a = 0.0
    if isinstance(alpha, torch.Tensor):
        alpha_rows = _need(alpha, torch.float32, "alpha", 1)
        if alpha_rows.shape[0] != n:
            raise ValueError("per-row alpha must have one entry per row")
    else:
        a = float(alpha)
    out = torch.empty_like(collab)
    d_pad = shadow_width(d)
    shadow = torch.empty((n, d_pad), dtype=torch.bfloat16, device=collab.device) if want_shadow else None
    with torch.cuda.device(collab.device):
        N.check(N.lib().hwer_blend_normalize(_dev_ptr(content), _dev_ptr(collab), a, _dev_ptr(alpha_rows), n, d,
                                             _dev_ptr(out), _dev_ptr(shadow), d_pad, _stream(collab.device)))
    return out, shadow


def unit_length(a):
    """hwer/utils.py:43-44 with axis=1 on a CUDA table."""
    return blend_normalize(None, a, 0.0, want_shadow=False)[0]


def make_shadow(table):
    table = _need(table, torch.float32, "table", 2)
    n, d = table.shape
    d_pad = shadow_width(d)
    shadow = torch.empty((n, d_pad), dtype=torch.bfloat16, device=table.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_make_shadow(_dev_ptr(table), n, d, _dev_ptr(shadow), d_pad, _stream(table.device)))
    return shadow


def norm_stats(table, epsilon=1e-4):
    """(violations, mean |norm-1|, positive, negative, max_norm) -- hwer/utils.py:51-57 plus the max norm."""
    table = _need(table, torch.float32, "table", 2)
    out = torch.empty(5, dtype=torch.float64, device=table.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_norm_stats(_dev_ptr(table), table.shape[0], table.shape[1], float(epsilon),
                                        _dev_ptr(out), _stream(table.device)))
    v = out.cpu().tolist()
    return int(v[0]), v[1], int(v[2]), int(v[3]), v[4]


class TopKIndex:
    """Exact top-k index over one [N, d] unit-norm fp32 table (one node type's rows): the replacement for a
    per-type `KDTree(vectors[rows], leaf_size=128)` of MultiKNN (hwer/recommendation_base.py:65-76)."""

    def __init__(self, table, shadow=None, max_norm=None):
        self.table = _need(table, torch.float32, "table", 2)
        self.n, self.d = self.table.shape
        self.device = self.table.device
        if shadow is None and shadow_width(self.d) <= 256:
            shadow = make_shadow(self.table)
        if shadow is not None:
            shadow = _need(shadow, torch.bfloat16, "shadow", 2)
            if shadow.shape[0] != self.n:
                raise ValueError("shadow and table row counts differ")
        self.shadow = shadow
        if max_norm is None:
            max_norm = norm_stats(self.table)[4]
        self.max_norm = float(max_norm)
        handle = c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_index_create(ctypes.byref(handle), _dev_ptr(self.table), _dev_ptr(shadow), self.n,
                                              self.d, shadow.shape[1] if shadow is not None else 0, self.max_norm,
                                              self.device.index if self.device.index is not None else
                                              torch.cuda.current_device()))
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                N.lib().hwer_index_destroy(h)
            except Exception:
                pass

    def topk_async(self, queries, k, mode="exact", idx_offset=0, cap=0, out=None, want_f64=False):
        """Enqueues the search on the current stream; results are valid after `finish()`."""
        queries = _need(queries, torch.float32, "queries", 2)
        if queries.shape[1] != self.d:
            raise ValueError("query width %d != table width %d" % (queries.shape[1], self.d))
        B = queries.shape[0]
        k = int(k)
        if k > self.n:
            # sklearn: "k must be less than or equal to the number of training points"
            raise ValueError("k=%d must be less than or equal to the number of rows %d" % (k, self.n))
        if out is None:
            idx = torch.empty((B, k), dtype=torch.int64, device=self.device)
            score = torch.empty((B, k), dtype=torch.float32, device=self.device)
            s64 = torch.empty((B, k), dtype=torch.float64, device=self.device) if want_f64 else None
        else:
            idx, score, s64 = out
        m = N.MODE_EXACT if mode == "exact" else N.MODE_BF16 if mode == "bf16" else None
        if m is None:
            raise ValueError("mode must be 'exact' or 'bf16'")
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_topk(self._h, _dev_ptr(queries), B, k, m, int(cap), int(idx_offset), _dev_ptr(idx),
                                      _dev_ptr(score), _dev_ptr(s64), _stream(self.device)))
        return idx, score, s64

    def topk_sharded_async(self, exchange, queries, k, mode="exact", idx_offset=0, cap=0, want_f64=False,
                           phases=N.PHASE_ALL, out=None):
        """hwer_topk_sharded: this shard's search, the peer-memory exchange and the owner-side merge, enqueued on
        the current stream.  Collective over the exchange's ranks; returns the global [B, k] result -- or, with
        PHASE_OWNED among `phases`, only the rows of the queries this rank merged (owner_range of the exchange)."""
        queries = _need(queries, torch.float32, "queries", 2)
        if queries.shape[1] != self.d:
            raise ValueError("query width %d != table width %d" % (queries.shape[1], self.d))
        B, k = queries.shape[0], int(k)
        if k > self.n:
            raise ValueError("k=%d must be less than or equal to the number of rows %d of every shard" % (k, self.n))
        if out is None:
            rows = B
            if int(phases) & N.PHASE_OWNED:
                lo, hi = exchange.owner_range(B)
                rows = hi - lo
            idx = torch.empty((rows, k), dtype=torch.int64, device=self.device)
            score = torch.empty((rows, k), dtype=torch.float32, device=self.device)
            s64 = torch.empty((rows, k), dtype=torch.float64, device=self.device) if want_f64 else None
        else:
            idx, score, s64 = out
        m = N.MODE_EXACT if mode == "exact" else N.MODE_BF16 if mode == "bf16" else None
        if m is None:
            raise ValueError("mode must be 'exact' or 'bf16'")
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_topk_sharded(self._h, exchange._h, _dev_ptr(queries), B, k, m, int(cap),
                                              int(idx_offset), _dev_ptr(idx), _dev_ptr(score), _dev_ptr(s64),
                                              int(phases), _stream(self.device)))
        return idx, score, s64

    def finish(self):
        need = c_uint32(0)
        with torch.cuda.device(self.device):
            rc = N.lib().hwer_topk_finish(self._h, _stream(self.device), ctypes.byref(need))
        return rc, int(need.value)

    MAX_CAP = 16384      # longest candidate list the selector can hold (csrc/api.cu make_schedule)

    def topk_exhaustive(self, queries, k, idx_offset=0):
        """hwer_topk_exhaustive: the same exact answer by scoring every row (slow; the path of last resort)."""
        queries = _need(queries, torch.float32, "queries", 2)
        B, k = queries.shape[0], int(k)
        idx = torch.empty((B, k), dtype=torch.int64, device=self.device)
        score = torch.empty((B, k), dtype=torch.float32, device=self.device)
        s64 = torch.empty((B, k), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_topk_exhaustive(self._h, _dev_ptr(queries), B, k, int(idx_offset), _dev_ptr(idx),
                                                 _dev_ptr(score), _dev_ptr(s64), _stream(self.device)))
        return idx, score, s64

    def topk(self, queries, k, mode="exact", idx_offset=0, want_f64=False):
        """Synchronous search.  Candidate-list overflow (heavily tied data) is retried with the capacity the kernels
        ask for; queries that need more than the selector can hold (MAX_CAP rows inside the bf16 margin of their
        k-th score, e.g. tens of thousands of duplicated cold-start embeddings) are answered exhaustively, so like
        the reference's KDTree the call always returns the exact result."""
        cap = 0
        for _ in range(6):
            idx, score, s64 = self.topk_async(queries, k, mode, idx_offset, cap, want_f64=True)
            rc, need = self.finish()
            if rc == N.HWER_OK:
                return (idx, score, s64) if want_f64 else (idx, score)
            if rc != N.HWER_E_OVERFLOW:
                N.check(rc)
            if need > self.MAX_CAP:
                marked = torch.nonzero(idx[:, 0] == -2).reshape(-1)          # rare path: torch glue is fine here
                if marked.numel():
                    e_idx, e_sc, e_s64 = self.topk_exhaustive(queries.index_select(0, marked).contiguous(), k, idx_offset)
                    idx[marked], score[marked], s64[marked] = e_idx, e_sc, e_s64
                return (idx, score, s64) if want_f64 else (idx, score)
            cap = need
        N.check(rc)

    def profile(self, enable=True):
        N.check(N.lib().hwer_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """(filter-kernel ms, filter launches, other kernel launches) since the last read; synchronises."""
        ms, fl, ol = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_profile_read(self._h, _stream(self.device), ctypes.byref(ms), ctypes.byref(fl),
                                              ctypes.byref(ol)))
        return ms.value, fl.value, ol.value

    STAGES = ("filter", "select", "final", "exchange")

    def profile_stages(self):
        """{stage: summed ms} of the bracketed launches since the last profile_read (call before it)."""
        buf = (ctypes.c_double * 4)()
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_profile_stages(self._h, _stream(self.device), buf))
        return dict(zip(self.STAGES, [buf[i] for i in range(4)]))

    def profile_launches(self, cap=4096):
        """Durations (ms) of the score-filter launches since the last profile_read, in launch order."""
        buf = (ctypes.c_double * cap)()
        n = ctypes.c_int32()
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_profile_launches(self._h, _stream(self.device), buf, cap, ctypes.byref(n)))
        return [buf[i] for i in range(min(n.value, cap))]

    def debug_scores(self, queries):
        queries = _need(queries, torch.float32, "queries", 2)
        out = torch.zeros((self.n, queries.shape[0]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().hwer_debug_scores(self._h, _dev_ptr(queries), queries.shape[0], _dev_ptr(out),
                                              queries.shape[0], _stream(self.device)))
        return out


def merge_topk(scores64, idx, want_f64=False):
    """[G, B, k] per-shard results -> [B, k] under (score desc, row asc)."""
    scores64 = _need(scores64, torch.float64, "scores64", 3)
    idx = _need(idx, torch.int64, "idx", 3)
    G, B, k = scores64.shape
    o_idx = torch.empty((B, k), dtype=torch.int64, device=idx.device)
    o_sc = torch.empty((B, k), dtype=torch.float32, device=idx.device)
    o_s64 = torch.empty((B, k), dtype=torch.float64, device=idx.device) if want_f64 else None
    with torch.cuda.device(idx.device):
        N.check(N.lib().hwer_merge_topk(_dev_ptr(scores64), _dev_ptr(idx), G, B, k, _dev_ptr(o_idx), _dev_ptr(o_sc),
                                        _dev_ptr(o_s64), _stream(idx.device)))
    return (o_idx, o_sc, o_s64) if want_f64 else (o_idx, o_sc)


def compose_queries(table, anchor_rows, pos=None, neg=None):
    """Query vectors of a batch of find_closest_neighbours calls (hwer/recommendation_base.py:164-170):
    average of unit(anchor row), unit(mean(positive rows)), -unit(mean(negative rows)) over the parts present.
    pos / neg: None or (ptr [B+1] int64, rows int64) CSR lists; row -1 = node never trained on."""
    table = _need(table, torch.float32, "table", 2)
    anchor_rows = _need(anchor_rows, torch.int64, "anchor_rows", 1)
    B = anchor_rows.shape[0]
    out = torch.empty((B, table.shape[1]), dtype=torch.float32, device=table.device)
    pp = pr = np_ = nr = None
    if pos is not None:
        pp, pr = _need(pos[0], torch.int64, "pos ptr", 1), _need(pos[1], torch.int64, "pos rows", 1)
        assert pp.shape[0] == B + 1
    if neg is not None:
        np_, nr = _need(neg[0], torch.int64, "neg ptr", 1), _need(neg[1], torch.int64, "neg rows", 1)
        assert np_.shape[0] == B + 1
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_compose_queries(_dev_ptr(table), table.shape[0], table.shape[1], _dev_ptr(anchor_rows),
                                             _dev_ptr(pp), _dev_ptr(pr), _dev_ptr(np_), _dev_ptr(nr), B, _dev_ptr(out),
                                             _stream(table.device)))
    return out


def average_embeddings(table, ptr, rows):
    """unit(mean(rows of each CSR list)) -- get_average_embeddings, hwer/recommendation_base.py:153-155.
    ptr [L+1] / rows int64 CUDA tensors; row -1 = node never trained on.  Returns [L, d] fp32."""
    table = _need(table, torch.float32, "table", 2)
    ptr = _need(ptr, torch.int64, "ptr", 1)
    rows = _need(rows, torch.int64, "rows", 1)
    L = ptr.shape[0] - 1
    out = torch.empty((L, table.shape[1]), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_average_embeddings(_dev_ptr(table), table.shape[0], table.shape[1], _dev_ptr(ptr),
                                                _dev_ptr(rows), L, _dev_ptr(out), _stream(table.device)))
    return out


def gather_rows(table, rows):
    """get_embeddings (hwer/recommendation_base.py:146-151): table[rows], clip(table[0], 1e-6, 1e-5) for row -1."""
    table = _need(table, torch.float32, "table", 2)
    rows = _need(rows, torch.int64, "rows", 1)
    out = torch.empty((rows.shape[0], table.shape[1]), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_gather_rows(_dev_ptr(table), table.shape[0], table.shape[1], _dev_ptr(rows), rows.shape[0],
                                         _dev_ptr(out), _stream(table.device)))
    return out


def map_rows(rows, row_map=None, offset=0):
    """Local rows of a per-type index -> global rows (negative ids pass through)."""
    rows = _need(rows, torch.int64, "rows")
    if row_map is not None:
        row_map = _need(row_map, torch.int64, "row_map", 1)
    out = torch.empty_like(rows)
    with torch.cuda.device(rows.device):
        N.check(N.lib().hwer_map_rows(_dev_ptr(rows), rows.numel(), _dev_ptr(row_map), int(offset), _dev_ptr(out),
                                      _stream(rows.device)))
    return out


_CONVENTIONS = {"pair": N.SCORE_PAIR, "dist": N.SCORE_DIST, "given": N.SCORE_GIVEN, "euclid": N.SCORE_EUCLID}


def rerank(table, rows, convention, anchor_rows=None, queries=None, given=None, row_map=None):
    """Scores the [B, k] retrieved rows in one of the reference's conventions and orders every anchor's list by
    that score (stable, like Python's sorted): "pair" (anchor . row + 1) / 2, "dist" (2 - ||row - query||) / 2,
    "given" caller-supplied scores -- all descending -- or "euclid" ||row - query|| ascending.  See hwer_rerank in
    include/hwer_b200.h.  Returns (rows [B, k] int64, scores [B, k] float64)."""
    table = _need(table, torch.float32, "table", 2)
    rows = _need(rows, torch.int64, "rows", 2)
    B, k = rows.shape
    if anchor_rows is not None:
        anchor_rows = _need(anchor_rows, torch.int64, "anchor_rows", 1)
    if queries is not None:
        queries = _need(queries, torch.float32, "queries", 2)
        if tuple(queries.shape) != (B, table.shape[1]):
            raise ValueError("queries must be [B, d]")
    if given is not None:
        given = _need(given, torch.float32, "given", 2)
        if tuple(given.shape) != (B, k):
            raise ValueError("given scores must be [B, k]")
    if row_map is not None:
        row_map = _need(row_map, torch.int64, "row_map", 1)
    out_rows = torch.empty_like(rows)
    out_score = torch.empty((B, k), dtype=torch.float64, device=rows.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_rerank(_dev_ptr(table), table.shape[0], table.shape[1], _dev_ptr(rows), _dev_ptr(row_map),
                                    B, k, _CONVENTIONS[convention], _dev_ptr(anchor_rows), _dev_ptr(queries),
                                    _dev_ptr(given), _dev_ptr(out_rows), _dev_ptr(out_score), _stream(table.device)))
    return out_rows, out_score


def hit_rank_metrics(scores, topn=10, want_rank=False):
    """HR@topn / binary NDCG@topn of the positive in column 0 of scores [U, 1 + M] (ncf_eval,
    hwer/validation.py:82-96).  Returns a float64 tensor {hr, ndcg} (and the [U] int32 ranks)."""
    scores = _need(scores, torch.float32, "scores", 2)
    U, M1 = scores.shape
    out = torch.empty(2, dtype=torch.float64, device=scores.device)
    rank = torch.empty(U, dtype=torch.int32, device=scores.device) if want_rank else None
    with torch.cuda.device(scores.device):
        N.check(N.lib().hwer_hit_rank_metrics(_dev_ptr(scores), U, M1 - 1, int(topn), _dev_ptr(out), _dev_ptr(rank),
                                              _stream(scores.device)))
    return (out, rank) if want_rank else out


def ncf_param_count(F, depth):
    return int(N.lib().hwer_ncf_param_count(int(F), int(depth)))


def ncf_score(h, params, src_rows, dst_rows, depth):
    """sigmoid(w_out . MLP([h[src] || h[dst]]) + b_out) per pair (hwer/ncf.py:7-27).  `h` is the reference's
    prediction_artifacts["h"] ([N + 1, F], row 0 = padding node); rows are node rows + 1, anything out of range
    reads row 0.  `params`: flat fp32 [W1, b1, ..., W_depth, b_depth, w_out, b_out] (torch Linear layout)."""
    h = _need(h, torch.float32, "h", 2)
    params = _need(params, torch.float32, "params", 1)
    src_rows = _need(src_rows, torch.int64, "src_rows", 1)
    dst_rows = _need(dst_rows, torch.int64, "dst_rows", 1)
    F = h.shape[1]
    if params.shape[0] != ncf_param_count(F, depth):
        raise ValueError("params has %d floats, an NCF of width %d and depth %d needs %d"
                         % (params.shape[0], F, depth, ncf_param_count(F, depth)))
    P = src_rows.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        N.check(N.lib().hwer_ncf_score(_dev_ptr(h), h.shape[0], F, int(depth), _dev_ptr(params), _dev_ptr(src_rows),
                                       _dev_ptr(dst_rows), P, _dev_ptr(out), _stream(h.device)))
    return out


def sample_neighbours(n, src, dst, fanout=2, seed=0, blocks=1):
    """Host-side stand-in for the reference's DGL NeighborSampler(g, batch, 2, layers, add_self_loop=True)
    (hwer/gcn_ncf.py:262-272): for each of `blocks` layers, up to `fanout` distinct random in-neighbours of every node
    over the undirected edge list (src[i], dst[i]) plus a self loop, as CSR (ptr [n + 1], idx) int64 numpy arrays.
    Seeded and vectorised (one random key per edge endpoint, the `fanout` smallest keys of every node win)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    to = np.concatenate([dst, src])                      # in-neighbour lists of both directions (gcn.py:206-215)
    frm = np.concatenate([src, dst])
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(int(blocks)):
        key = rs.random_sample(to.shape[0])
        order = np.lexsort((key, to))                    # by node, then by random key
        to_s, frm_s = to[order], frm[order]
        start = np.searchsorted(to_s, np.arange(n))      # first position of every node's run
        rank = np.arange(to_s.shape[0]) - start[to_s]
        keep = rank < fanout
        picked_to, picked_from = to_s[keep], frm_s[keep]
        all_to = np.concatenate([picked_to, np.arange(n, dtype=np.int64)])      # + self loops
        all_from = np.concatenate([picked_from, np.arange(n, dtype=np.int64)])
        o2 = np.argsort(all_to, kind="stable")
        idx = all_from[o2]
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(all_to, minlength=n), out=ptr[1:])
        out.append((ptr, idx))
    return out


def gcn_infer(node_emb, content, proj_w, proj_b, ln_g, ln_b, nbr, fc0_w, fc0_b, fc1_w, fc1_b, previous=None, ema=0.1):
    """GraphConvModule.forward in eval mode over the whole graph (hwer/gcn.py:162-193; get_gcn_vectors,
    hwer/gcn_ncf.py:260-279) with the neighbour sample of every block given as CSR lists `nbr[i] = (ptr, idx)`.
    All tensors on one CUDA device, fp32 (lists int64); `previous` ([>= n, F]) is updated in place.
    Returns h [n, F].  Widths that are not a multiple of 4 are zero-padded here (the result is unchanged)."""
    node_emb = _need(node_emb, torch.float32, "node_emb", 2)
    content = _need(content, torch.float32, "content", 2)
    n, C = content.shape
    F = node_emb.shape[1]
    layers = len(nbr)
    dev = content.device
    proj_w = _need(proj_w, torch.float32, "proj_w", 2)
    fc0_w = _need(fc0_w, torch.float32, "fc0_w", 2)
    fc1_w = _need(fc1_w, torch.float32, "fc1_w", 2)
    if node_emb.shape[0] < n + 1 or proj_w.shape != (F, C) or fc0_w.shape != (4 * F, F * (layers + 1)) or \
            fc1_w.shape != (F, 4 * F):
        raise ValueError("gcn_infer: parameter shapes do not match n=%d, C=%d, F=%d, layers=%d" % (n, C, F, layers))
    if F % 4:
        raise ValueError("gcn_infer: the feature width must be a multiple of 4 (the reference asserts a multiple of 16)")
    if C % 4:                                            # zero columns of content / proj_w do not change the product
        pad = 4 - C % 4
        content = torch.nn.functional.pad(content, (0, pad)).contiguous()
        proj_w = torch.nn.functional.pad(proj_w, (0, pad)).contiguous()
        C += pad
    vecs = [_need(t, torch.float32, name, 1) for t, name in ((proj_b, "proj_b"), (ln_g, "ln_g"), (ln_b, "ln_b"),
                                                             (fc0_b, "fc0_b"), (fc1_b, "fc1_b"))]
    ptrs = [_need(p, torch.int64, "nbr ptr", 1) for p, _ in nbr]
    idxs = [_need(i, torch.int64, "nbr idx", 1) for _, i in nbr]
    if any(p.shape[0] != n + 1 for p in ptrs):
        raise ValueError("gcn_infer: every neighbour list needs n + 1 offsets")
    if previous is not None:
        previous = _need(previous, torch.float32, "previous", 2)
        if previous.shape[0] < n or previous.shape[1] != F:
            raise ValueError("gcn_infer: previous must be [>= n, F]")
    out = torch.empty((n, F), dtype=torch.float32, device=dev)
    p_arr = (ctypes.c_void_p * layers)(*[_dev_ptr(p) for p in ptrs])
    i_arr = (ctypes.c_void_p * layers)(*[_dev_ptr(i) for i in idxs])
    with torch.cuda.device(dev):
        N.check(N.lib().hwer_gcn_infer(_dev_ptr(node_emb), _dev_ptr(content), n, C, F, layers, _dev_ptr(proj_w),
                                       _dev_ptr(vecs[0]), _dev_ptr(vecs[1]), _dev_ptr(vecs[2]), p_arr, i_arr,
                                       _dev_ptr(fc0_w), _dev_ptr(vecs[3]), _dev_ptr(fc1_w), _dev_ptr(vecs[4]),
                                       _dev_ptr(previous), float(ema), _dev_ptr(out), _stream(dev)))
    return out


def pair_score(table, src_rows, dst_rows):
    """(dot + 1) / 2 of row pairs; row -1 = node unseen in training (hwer/recommendation_base.py:135-151)."""
    table = _need(table, torch.float32, "table", 2)
    src_rows = _need(src_rows, torch.int64, "src_rows", 1)
    dst_rows = _need(dst_rows, torch.int64, "dst_rows", 1)
    P = src_rows.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        N.check(N.lib().hwer_pair_score(_dev_ptr(table), table.shape[0], table.shape[1], _dev_ptr(src_rows),
                                        _dev_ptr(dst_rows), P, _dev_ptr(out), _stream(table.device)))
    return out


def eval_metrics(topk_items, train_ptr, train_idx, val_ptr, val_idx, val_rel, cutoffs, n_items, per_user=False):
    """Ranking metrics of hwer/validation.py:133-174 for all users at once; see include/hwer_b200.h for layouts."""
    topk_items = _need(topk_items, torch.int64, "topk_items", 2)
    dev = topk_items.device
    U, kret = topk_items.shape
    cut = torch.tensor(sorted(int(c) for c in cutoffs), dtype=torch.int32, device=dev)
    if cut[-1].item() > 256 or cut[0].item() <= 0:
        raise ValueError("cutoffs must lie in [1, 256]")
    n_cut = cut.shape[0]
    out = torch.empty(3 * n_cut + 3, dtype=torch.float64, device=dev)
    pu = torch.empty((U, 3 * n_cut + 1), dtype=torch.float64, device=dev) if per_user else None
    for nm, t in (("train_ptr", train_ptr), ("train_idx", train_idx), ("val_ptr", val_ptr), ("val_idx", val_idx)):
        _need(t, torch.int64, nm, 1)
    _need(val_rel, torch.float32, "val_rel", 1)
    with torch.cuda.device(dev):
        N.check(N.lib().hwer_eval_metrics(_dev_ptr(topk_items), U, kret, _dev_ptr(train_ptr), _dev_ptr(train_idx),
                                          _dev_ptr(val_ptr), _dev_ptr(val_idx), _dev_ptr(val_rel), _dev_ptr(cut),
                                          n_cut, int(n_items), _dev_ptr(out), _dev_ptr(pu), _stream(dev)))
    return (out, pu) if per_user else out


LINK_METRIC_NAMES = ("ap", "precision", "recall", "accuracy", "tp", "fp", "fn", "tn")


def link_metrics(scores, labels, threshold=0.5):
    """Average precision and precision / recall / accuracy at `threshold` of scored 0/1-labelled pairs, on the
    device (the sklearn calls of hwer/validation.py:52-59).  Returns a float64 tensor of 8, see LINK_METRIC_NAMES."""
    scores = _need(scores, torch.float32, "scores", 1)
    labels = _need(labels, torch.uint8, "labels", 1)
    if labels.shape[0] != scores.shape[0] or scores.shape[0] == 0:
        raise ValueError("link_metrics: scores and labels must be non-empty and of equal length")
    out = torch.empty(8, dtype=torch.float64, device=scores.device)
    with torch.cuda.device(scores.device):
        N.check(N.lib().hwer_link_metrics(_dev_ptr(scores), _dev_ptr(labels), scores.shape[0], float(threshold),
                                          _dev_ptr(out), _stream(scores.device)))
    return out

"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + sklearn's KDTree) of the reference hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module, and only as the checker or the timed CPU baseline.  Nothing under
`hybrid-weighted-embedding-recommender_b200/` imports it; the product has no CPU path.

Every function cites the reference lines it restates (paths relative to the reference checkout).  The k-NN
arithmetic itself lives in a third-party dependency that is NOT under /root/reference:
`sklearn.neighbors.KDTree` (pinned scikit-learn==0.21.3 at requirements.txt:112; this image has 1.9.0, whose
KDTree == KDTree64 accepts the same calls).  Its published algorithm -- exact Euclidean k-NN in float64 -- is
what `exact_topk` restates by brute force; parity is anchored on the reference's own call sites
(recommendation_base.py:74,79).

Pinning: the reference ships no golden vectors or asserting tests for this path.
This oracle is instead pinned against outputs of the REFERENCE ITSELF executed in the build container
(`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).  The alpha blend
has no reference code at all (only README prose, README.md:2,84,110-114); `blend_normalize` restates
BASELINE.json's north_star formula and is pinned only at alpha = 0, where it must equal the reference's
`prepare_for_knn` (gcn_ncf.py:447-456) -- for 0 < alpha <= 1 parity is UNPINNED.
"""
import operator
from collections import defaultdict

import numpy as np
from numpy.linalg import norm
from sklearn.neighbors import KDTree


# --------------------------------------------------------------------------- utils.py
def unit_length(a, axis=0):
    """hwer/utils.py:43-44 -- a / ||a||, no epsilon (a zero row becomes NaN)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return a / np.expand_dims(norm(a, axis=axis), axis=axis)


def unit_length_violations(a, axis=0, epsilon=1e-4):
    """hwer/utils.py:51-57."""
    vector_lengths = np.expand_dims(norm(a, axis=axis), axis=axis)
    positive_violations = np.sum(vector_lengths > 1 + epsilon)
    negative_violations = np.sum(vector_lengths < 1 - epsilon)
    violations = positive_violations + negative_violations
    violation_mean = np.mean(np.abs(vector_lengths - 1))
    return violations, violation_mean, positive_violations, negative_violations


def blend_normalize(content, collab, alpha):
    """north_star blend in the slot of GcnNCF.prepare_for_knn (hwer/gcn_ncf.py:447-456):
    V = unit(alpha * unit(C) + (1 - alpha) * unit(G)) row-wise.  alpha may be a scalar or an [N] vector.
    At alpha = 0 this is exactly the reference's `unit_length(collaborative_vectors, axis=1)`."""
    collab = np.asarray(collab)
    g = unit_length(collab, axis=1)
    if content is None or (np.isscalar(alpha) and float(alpha) == 0.0):
        return g            # the reference's code path: a single normalisation, content ignored
    c = unit_length(np.asarray(content), axis=1)
    a = np.asarray(alpha, dtype=collab.dtype)
    if a.ndim == 1:
        a = a[:, None]
    return unit_length(a * c + (1 - a) * g, axis=1)


def prepare_for_knn(collaborative_vectors, n_dims):
    """hwer/gcn_ncf.py:447-456 (content_vectors is ignored there): PCA to n_dims when the table is wider -- sklearn's
    PCA, a third-party dependency outside /root/reference, restated as the exact decomposition: centre, SVD, project
    on the leading right singular vectors, each flipped so that its largest-magnitude loading is positive (sklearn's
    svd_flip rule) -- then unit_length.  Equal up to solver error to whatever solver sklearn's 'auto' picks."""
    x = np.asarray(collaborative_vectors)
    if x.shape[1] > n_dims:
        xc = x.astype(np.float64) - x.astype(np.float64).mean(axis=0)
        _, _, vt = np.linalg.svd(xc, full_matrices=False)
        vt = vt[:n_dims]
        vt = vt * np.sign(vt[np.arange(n_dims), np.abs(vt).argmax(axis=1)])[:, None]
        x = (xc @ vt.T).astype(np.float32)
    elif x.shape[1] < n_dims:
        raise ValueError()
    return unit_length(x, axis=1)


def reciprocal_rank(y_true, y_pred):
    """hwer/utils.py:71-78."""
    y_true = set(y_true)
    for i, e in enumerate(y_pred):
        if e in y_true:
            return 1.0 / (i + 1)
    return 0.0


def ndcg(y_true, y_pred):
    """hwer/utils.py:101-107 (gain 2^rel - 1, discount log2(i + 2), ideal list truncated to len(y_pred))."""
    y_true_sorted = sorted(y_true.values(), reverse=True)
    y_true_sorted = y_true_sorted[:len(y_pred)]
    idcg = np.sum((np.power(2, y_true_sorted) - 1) / (np.log2(np.arange(len(y_true_sorted)) + 2)))
    y_pred = [y_true[i] if i in y_true else 0 for i in y_pred]
    dcg = np.sum((np.power(2, y_pred) - 1) / (np.log2(np.arange(len(y_pred)) + 2)))
    return dcg / (idcg + 1e-8)


def binary_ndcg(y_true, y_pred):
    """hwer/utils.py:110-111."""
    return ndcg({k: 1 for k, v in y_true.items()}, y_pred)


def binary_ndcg_v2(y_true, y_pred):
    """hwer/utils.py:114-115."""
    return ndcg({k: 1 for k in y_true}, y_pred)


def recall(y_true, y_pred):
    """hwer/utils.py:118-121."""
    nrm = min(len(y_pred), len(y_true))
    hits = sum([1 if i in y_true else 0 for i in y_pred])
    return hits / max(nrm, 1.0)


# --------------------------------------------------------------------------- exact brute-force k-NN
def ncf_layer_dims(F, depth):
    """(in, out) of every Linear of the reference NCF, hwer/ncf.py:12-16, then the (F, 1) output layer :19."""
    dims = []
    for layer_idx in range(1, depth + 1):
        iw = 4 if layer_idx == 2 else 2
        ow = 1 if layer_idx == depth else (4 if layer_idx == 1 else 2)
        dims.append((F * iw, F * ow))
    return dims + [(F, 1)]


def ncf_forward(h, src, dst, params, depth):
    """hwer/ncf.py:24-27 at inference (GaussianNoise is the identity in eval mode, hwer/gcn.py:32-38): fp32
    Linear + LeakyReLU(0.01) stack over [h[src] || h[dst]], Linear(F, 1), sigmoid.  `params` = flat
    [W1 (out x in), b1, ..., w_out, b_out]."""
    h = np.asarray(h, dtype=np.float32)
    x = np.concatenate([h[src], h[dst]], axis=1)
    dims = ncf_layer_dims(h.shape[1], depth)
    off = 0
    for li, (i, o) in enumerate(dims):
        w = params[off:off + i * o].reshape(o, i).astype(np.float32); off += i * o
        b = params[off:off + o].astype(np.float32); off += o
        x = x @ w.T + b
        if li < len(dims) - 1:
            x = np.where(x > 0, x, np.float32(0.01) * x)
    assert off == len(params)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).reshape(-1)


def exact_topk(table, queries, k, block=4096):
    """Exact top-k by dot product: float64 scores over the stored rows, ordered (-score, row).
    This is what KDTree64.query computes on unit rows (Euclidean order == dot order),
    with the arbitrary KD-tree tie order replaced by the documented rule."""
    t64 = np.asarray(table, dtype=np.float64)
    q64 = np.asarray(queries, dtype=np.float64)
    n = t64.shape[0]
    k = min(k, n)
    idx = np.empty((q64.shape[0], k), dtype=np.int64)
    sc = np.empty((q64.shape[0], k), dtype=np.float64)
    for b0 in range(0, q64.shape[0], block):
        s = q64[b0:b0 + block] @ t64.T
        for r in range(s.shape[0]):
            row = s[r]
            if k < n:
                # everything tied with the k-th score must take part in the (score, row) ordering
                kth = np.partition(row, n - k)[n - k]
                cand = np.nonzero(row >= kth)[0]
            else:
                cand = np.arange(n)
            order = np.lexsort((cand, -row[cand]))[:k]
            idx[b0 + r] = cand[order]
            sc[b0 + r] = row[cand[order]]
    return idx, sc


def compare_topk(idx, score, ref_idx, ref_score, tie_eps=1e-6):
    """Tie-aware comparison: rows must agree position by position except inside groups of
    reference scores closer than `tie_eps`, where only the sets must agree (the group straddling the k-th place
    may swap members with equally-scored rows just outside).  Returns the number of mismatching queries."""
    bad = 0
    idx = np.asarray(idx)
    ref_idx = np.asarray(ref_idx)
    for r in range(ref_idx.shape[0]):
        if np.array_equal(idx[r], ref_idx[r]):
            continue
        k = ref_idx.shape[1]
        ok = True
        i = 0
        while i < k:
            j = i + 1
            while j < k and abs(ref_score[r, j] - ref_score[r, j - 1]) <= tie_eps:
                j += 1
            if set(idx[r, i:j]) != set(ref_idx[r, i:j]):
                # tolerated only when the group touches the cut and the scores match within tie_eps
                if j == k and np.all(np.abs(np.sort(score[r, i:j])[::-1] - ref_score[r, i:j]) <= tie_eps):
                    pass
                else:
                    ok = False
                    break
            i = j
        bad += 0 if ok else 1
    return bad


# --------------------------------------------------------------------------- recommendation_base.py
class Node:
    """hwer/recommendation_base.py:19-36."""

    def __init__(self, node_type, node_external_id):
        self.node_type = node_type
        self.node_external_id = str(node_external_id)

    def _key(self):
        return (self.node_type, self.node_external_id)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        if isinstance(other, Node):
            return self._key() == other._key()
        return NotImplemented

    def __repr__(self):
        return str(self._key())


class NodeNotFoundException(Exception):
    """hwer/utils.py:326."""


class MultiKNN:
    """hwer/recommendation_base.py:64-83 -- one exact KD-tree per node type over that type's rows."""

    def __init__(self, nodes_to_idx, vectors, leaf_size=128):
        self.nodes = [None] * len(nodes_to_idx)
        rows = defaultdict(list)
        for n, i in nodes_to_idx.items():
            self.nodes[i] = n
            rows[n.node_type].append(i)
        self.idxs = {k: np.asarray(v) for k, v in rows.items()}
        self.knn = {k: KDTree(vectors[v], leaf_size=leaf_size) for k, v in rows.items()}

    def query(self, embedding, node_type, k=200):
        (dist,), (neighbors,) = self.knn[node_type].query([embedding], k=k)
        rows = self.idxs[node_type][neighbors]
        results = [(self.nodes[i], dt) for i, dt in zip(rows, dist)]
        return list(sorted(results, key=operator.itemgetter(1), reverse=False))


class OracleRecommender:
    """hwer/recommendation_base.py:86-174 restated around an externally supplied table (the reference's
    subclasses produce the table in fit(); the serving methods below do not depend on how).
    `gcn_scores=True` selects the GcnNCF variant of the final scoring (hwer/gcn_ncf.py:378-383)."""

    def __init__(self, node_types, n_dims=32, gcn_scores=False):
        self.node_types = set(node_types)
        self.nodes_to_idx = {}
        self.knn = None
        self.vectors = None
        self.fit_done = False
        self.n_dims = n_dims
        self.gcn_scores = gcn_scores

    def add_nodes(self, nodes):                                    # :96-103
        assert len(set(nodes)) == len(nodes)
        assert self.nodes_to_idx.keys().isdisjoint(set(nodes))
        assert len(set([n.node_type for n in nodes]) - self.node_types) == 0
        base = len(self.nodes_to_idx)
        self.nodes_to_idx.update(zip(nodes, range(base, base + len(nodes))))
        return self

    def build_knn(self, vectors):                                  # :105-110
        v, _, _, _ = unit_length_violations(vectors, axis=1)
        assert v == 0
        self.knn = MultiKNN(self.nodes_to_idx, vectors, leaf_size=128)
        self.vectors = vectors
        self.fit_done = True
        return self

    def get_embeddings(self, nodes):                               # :146-151
        indexes = np.array([self.nodes_to_idx[n] if n in self.nodes_to_idx else -1 for n in nodes])
        mask = indexes == -1
        embeddings = self.vectors[np.where(indexes >= 0, indexes, 0)]
        embeddings[mask] = np.clip(embeddings[mask], 1e-6, 1e-5)
        return embeddings

    def get_average_embeddings(self, entities):                    # :153-155
        return unit_length(np.average(self.get_embeddings(entities), axis=0))

    def predict(self, node_pairs):                                 # :135-144 (== gcn_ncf.py:330-334)
        src, dst = zip(*node_pairs)
        results = (self.get_embeddings(src) * self.get_embeddings(dst)).sum(1)
        return (results + 1) / 2

    def query_embedding(self, anchor, positive=None, negative=None):   # :164-170
        embedding_list = [self.get_average_embeddings([anchor])]
        if positive is not None and len(positive) > 0:
            embedding_list.append(self.get_average_embeddings(positive))
        if negative is not None and len(negative) > 0:
            embedding_list.append(-1 * self.get_average_embeddings(negative))
        return np.average(embedding_list, axis=0)

    def find_closest_neighbours(self, node_type, anchor, positive=None, negative=None, k=200):   # :157-174
        assert self.fit_done
        assert node_type in self.node_types and node_type in self.knn.knn
        if anchor not in self.nodes_to_idx:
            raise NodeNotFoundException("Node = %s, was not provided in training" % anchor)
        embedding = self.query_embedding(anchor, positive, negative)
        node_dist_list = self.knn.query(embedding, node_type, k=k)
        if self.gcn_scores:                                        # gcn_ncf.py:378-383
            nodes, dist = zip(*node_dist_list)
            dist = (-1 * np.array(dist) + 2) / 2
            return list(sorted(zip(nodes, dist), key=operator.itemgetter(1), reverse=True))
        scores = self.predict([(anchor, node) for node, dist in node_dist_list])
        return list(sorted(zip([n for n, d in node_dist_list], scores), key=operator.itemgetter(1), reverse=True))


def model_get_topk_knn(model, anchors, node_type, k=200):
    """hwer/validation.py:30-35 -- the serial per-anchor loop (k added: the reference uses the default 200)."""
    predictions = defaultdict(list)
    for u in anchors:
        predictions[u] = model.find_closest_neighbours(node_type, u, k=k)
    return predictions


# --------------------------------------------------------------------------- validation.py
def extraction_metrics(predictions, train_edges, validation_edges, node_type, cutoffs=(10, 20, 50, 100)):
    """hwer/validation.py:100-187 minus the model calls: `predictions` is what get_topk returned
    ({user: [(item, score), ...]}).  Returns per-cutoff recall / ndcg / binary ndcg, MRR@max and diversity."""
    validation_users = list(set([u for u, i, r in validation_edges]))
    train_items = list(set([i for u, i, r in train_edges]))
    validation_items = list(set([i for u, i, r in validation_edges]))
    all_items = [x for x in set(validation_items + train_items) if x.node_type == node_type]
    train_uid = defaultdict(set)
    for u, i, r in train_edges:
        train_uid[u].add(i)
    filtered = {}
    for u, i in predictions.items():                               # :133-141
        remaining = [it for it, r in sorted(i, key=operator.itemgetter(1), reverse=True)]
        filtered[u] = [x for x in remaining if x not in train_uid[u]]
    validation_actuals = defaultdict(list)
    for u, i, r in validation_edges:
        validation_actuals[u].append((i, r))
    score_dict = defaultdict(dict)
    for u, i in validation_actuals.items():                        # :156-163
        remaining = sorted(i, key=operator.itemgetter(1), reverse=True)
        remaining = [x for x in remaining if x[0] not in train_uid[u]]
        score_dict[u] = dict(remaining)
        validation_actuals[u] = [it for it, r in remaining]
    out = {}
    for c in cutoffs:
        p = {u: filtered.get(u, [])[:c] for u in validation_users}
        out["recall@%d" % c] = np.mean([recall(score_dict[u], p[u]) for u in validation_users])
        out["ndcg@%d" % c] = np.mean([ndcg(score_dict[u], p[u]) for u in validation_users])
        out["ndcg_b@%d" % c] = np.mean([binary_ndcg(score_dict[u], p[u]) for u in validation_users])
    top = max(cutoffs)
    out["mrr"] = np.mean([reciprocal_rank(validation_actuals[u], filtered.get(u, [])[:top]) for u in validation_users])
    seen = set()
    for u, p in filtered.items():
        seen.update(p[:top])
    out["diversity"] = len(seen) / max(len(all_items), 1)          # :144-145
    return out


def average_precision_score(labels, scores):
    """sklearn.metrics.average_precision_score for binary labels (the call at hwer/validation.py:52-53; sklearn is a
    third-party dependency outside /root/reference, pinned scikit-learn==0.21.3): AP = sum_n (R_n - R_{n-1}) P_n over
    the DISTINCT score thresholds, scores sorted descending.  Restated with numpy, float64."""
    labels = np.asarray(labels).astype(np.float64)
    scores = np.asarray(scores)
    order = np.argsort(-scores, kind="mergesort")
    s, y = scores[order], labels[order]
    ends = np.r_[np.where(np.diff(s))[0], len(s) - 1]             # last index of each group of equal scores
    tps = np.cumsum(y)[ends]
    precision = tps / (ends + 1.0)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def link_prediction_metrics(labels, predictions, threshold=0.5):
    """The metric half of hwer/validation.py:52-64: AP plus precision / recall (average='binary') and accuracy of
    `predictions >= 0.5`.  Returns (ap, precision, recall, accuracy)."""
    labels = np.asarray(labels).astype(bool)
    pred = np.asarray(predictions) >= threshold
    tp = float(np.sum(labels & pred)); fp = float(np.sum(~labels & pred))
    fn = float(np.sum(labels & ~pred)); tn = float(np.sum(~labels & ~pred))
    precision = tp / (tp + fp) if tp + fp > 0 else 0.0
    rec = tp / (tp + fn) if tp + fn > 0 else 0.0
    return average_precision_score(labels, predictions), precision, rec, (tp + tn) / max(len(labels), 1)


def link_prediction_accuracy(model, nodes, train_edges, validation_edges, rng):
    """hwer/validation.py:41-65.  `rng` is the `random` module (or a random.Random): the reference draws
    10 x |E| negative pairs with random.choices in this exact order -- train src, train dst, validation src,
    validation dst -- so a seeded generator reproduces its pair sets."""
    m = 10
    out = {}
    sets = {}
    for name, edges in (("train", train_edges), ("val", validation_edges)):
        neg_src = rng.choices(nodes, k=len(edges) * m)
        neg_dst = rng.choices(nodes, k=len(edges) * m)
        sets[name] = ([(u, i) for u, i, r in edges] + list(zip(neg_src, neg_dst)),
                      [1] * len(edges) + [0] * (len(edges) * m))
    for name in ("train", "val"):                                  # predictions after BOTH sets are drawn (:44-51)
        pairs, labels = sets[name]
        ap, p, r, a = link_prediction_metrics(labels, np.array(model.predict(pairs)))
        out["lp_%s_ap" % name], out["lp_%s_precision" % name] = ap, p
        out["lp_%s_recall" % name], out["lp_%s_accuracy" % name] = r, a
    return out


def ncf_eval(model, train_edges, validation_edges, item_list, rng):
    """hwer/validation.py:68-97: per validation edge 1 positive + 100 negatives drawn with random.sample from the
    items the user never touched, scored by model.predict, stable descending sort, top 10 -> HR@10 and
    binary_ndcg_v2([positive], top10), averaged over users (one entry per user: the last edge wins, :79-81).
    Python >= 3.11 refuses to sample from a set, so the pool is given the one deterministic order a set of Nodes
    has -- sorted by repr -- exactly what oracle/ref_shim.py does to run the unmodified reference here.
    `rng`: the `random` module or a random.Random.  Returns (ncf_hr, ncf_ndcg, {user: rank of the positive})."""
    item_list = set(item_list)
    interactions = defaultdict(set)
    for u, i, _ in train_edges:
        interactions[u].add(i)
    for u, i, _ in validation_edges:
        interactions[u].add(i)
    user_test_item, actual = {}, {}
    for u, i, _ in validation_edges:
        user_test_item[u] = [i, *rng.sample(sorted(item_list - interactions[u], key=repr), 100)]
        actual[u] = i
    top10, ranks = {}, {}
    for u, items in user_test_item.items():
        it = list(zip(items, model.predict([(u, i) for i in items])))
        it = list(sorted(it, key=operator.itemgetter(1), reverse=True))
        ranks[u] = [x for x, _ in it].index(items[0])
        top10[u], _ = zip(*it[:10])
    hr = [actual[u] in top10[u] for u in actual]
    nd = [binary_ndcg_v2([actual[u]], top10[u]) for u in actual]
    return float(np.mean(hr)), float(np.mean(nd)), ranks


# ------------------------------------------------------------------------------------------- GCN inference (8f rank 4)
def gcn_infer(node_emb, content, proj_w, proj_b, ln_g, ln_b, nbr, fc0_w, fc0_b, fc1_w, fc1_b, previous=None, ema=0.1):
    """Restatement of GraphConvModule.forward in eval mode (hwer/gcn.py:162-193) over the whole graph, with the
    neighbour sample of every block given explicitly (`nbr[i] = (ptr [n+1], idx)`, self loop included) -- the form
    get_gcn_vectors (hwer/gcn_ncf.py:260-279) takes when the NodeFlow covers every node.  fp32 like the reference.

      h0[v]   = unit(node_emb[v + 1] + LayerNorm(LeakyReLU_0.1(content[v] W^T + b)))          gcn.py:40-44,59-63,170-175
      H_i[v]  = [ mean_{u in nbr_i(v)} H_{i-1}[u]  ||  h0[v] ]                                 gcn.py:119-125,159-160
      out[v]  = unit(fc1(LeakyReLU_0.01(fc0(H_L[v]))))                                         gcn.py:104-114,126-128
      out[v]  = (1 - ema) out[v] + ema previous[v];  previous[v] = out[v]                      gcn.py:186-191
    (unit() divides by max(norm, 1e-5); GaussianNoise is the identity in eval mode, gcn.py:32-37.)
    Returns (out [n, F], previous_after or None)."""
    f32 = np.float32
    node_emb = np.asarray(node_emb, f32)
    content = np.asarray(content, f32)
    n = content.shape[0]

    def unit(a):
        nrm = np.sqrt((a * a).sum(axis=1, keepdims=True, dtype=f32)).astype(f32)
        return (a / np.maximum(nrm, f32(1e-5))).astype(f32)

    c = content @ np.asarray(proj_w, f32).T + np.asarray(proj_b, f32)
    c = np.where(c > 0, c, f32(0.1) * c).astype(f32)
    mu = c.mean(axis=1, keepdims=True, dtype=f32)
    var = ((c - mu) ** 2).mean(axis=1, keepdims=True, dtype=f32)
    c = ((c - mu) / np.sqrt(var + f32(1e-5)) * np.asarray(ln_g, f32) + np.asarray(ln_b, f32)).astype(f32)
    h0 = unit(node_emb[1:n + 1] + c)
    H = h0
    for ptr, idx in nbr:
        ptr = np.asarray(ptr, np.int64)
        idx = np.asarray(idx, np.int64)
        agg = np.zeros_like(H)
        for v in range(n):
            rows = idx[ptr[v]:ptr[v + 1]]
            s = np.zeros(H.shape[1], f32)
            for u in rows:                                   # list order, like DGL's sum reducer
                s = s + H[u]
            agg[v] = s / f32(len(rows))
        H = np.concatenate([agg, h0], axis=1).astype(f32)
    z = H @ np.asarray(fc0_w, f32).T + np.asarray(fc0_b, f32)
    z = np.where(z > 0, z, f32(0.01) * z).astype(f32)
    z = z @ np.asarray(fc1_w, f32).T + np.asarray(fc1_b, f32)
    out = unit(z.astype(f32))
    prev_after = None
    if previous is not None:
        previous = np.asarray(previous, f32)
        out = (f32(1.0 - ema) * out + f32(ema) * previous[:n]).astype(f32)      # previous is indexed by node id, gcn.py:188
        prev_after = previous.copy()
        prev_after[:n] = out
    return out, prev_after

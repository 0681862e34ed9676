"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded inputs, against the
committed outputs of the reference itself (tests/golden), and -- at BASELINE.json's full size -- through
size-independent properties.  Tolerances: indices bit-exact apart from documented ties (|score gap| <= 1e-6);
scores within 1e-5 (exact mode) / 2e-2 relative (bf16 mode), as BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

import hwer_oracle as O
from conftest import posneg_lists, synthetic_case, synthetic_edges

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hw():
    import hwer_b200
    from hwer_b200 import _native
    assert _native.library_path().endswith("libhwer_b200.so")
    return hwer_b200


def unit_table(n, d, seed, device="cuda"):
    rs = np.random.RandomState(seed)
    t = O.unit_length(rs.standard_normal((n, d)).astype(np.float32), axis=1)
    return t, torch.from_numpy(t).to(device)


# ----------------------------------------------------------------------------- tensor-core scorer
@pytest.mark.parametrize("n,d,B", [(1000, 128, 5), (4097, 64, 48), (3000, 256, 130), (2500, 100, 17),
                                   (70000, 128, 300), (129, 192, 1)])
def test_tensor_core_scores_match_bf16_product(hw, n, d, B):
    """Every element of the TMEM score matrix (debug dump) == dot of the bf16-rounded operands."""
    t_np, t = unit_table(n, d, 1)
    q_np, q = unit_table(B, d, 2)
    index = hw.ops.TopKIndex(t)
    got = index.debug_scores(q).cpu().numpy()                       # [n, B]
    tb = t.to(torch.bfloat16).double()
    qb = q.to(torch.bfloat16).double()
    want = (tb @ qb.t()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    # and the proven bound that makes the exact mode exact: |bf16 score - exact score| <= eps * |q||x|
    exact = t_np.astype(np.float64) @ q_np.astype(np.float64).T
    assert np.abs(got - exact).max() <= 0.00390625 + 3.9e-6 + ((d + 63) // 64 * 64) * 2.4e-7


# ----------------------------------------------------------------------------- blend + normalise
@pytest.mark.parametrize("d", [64, 128, 100, 256, 7])
@pytest.mark.parametrize("alpha", [0.0, 0.25, 0.5, 1.0, "rows"])
def test_blend_normalize(hw, d, alpha):
    n = 3001
    rs = np.random.RandomState(3)
    c = rs.standard_normal((n, d)).astype(np.float32) * 3.0
    g = rs.standard_normal((n, d)).astype(np.float32) * 0.2
    a = rs.rand(n).astype(np.float32) if alpha == "rows" else alpha
    want = O.blend_normalize(c, g, a)
    a_dev = torch.from_numpy(a).cuda() if alpha == "rows" else alpha
    table, shadow = hw.ops.blend_normalize(torch.from_numpy(c).cuda(), torch.from_numpy(g).cuda(), a_dev)
    np.testing.assert_allclose(table.cpu().numpy(), want, rtol=0, atol=1e-6)       # fp tolerance: 1e-6
    assert shadow.shape == (n, (d + 63) // 64 * 64) and shadow.dtype == torch.bfloat16
    assert torch.equal(shadow[:, :d], table.to(torch.bfloat16))                     # round-to-nearest-even, bit exact
    assert bool((shadow[:, d:] == 0).all())
    v, mean, _, _, mx = hw.ops.norm_stats(table)
    assert v == 0 and mean < 1e-6 and abs(mx - 1) < 1e-5


def test_unit_length_and_zero_row(hw):
    rs = np.random.RandomState(4)
    a = rs.standard_normal((257, 128)).astype(np.float32)
    a[5] = 0.0
    want = O.unit_length(a, axis=1)
    got = hw.unit_length(a, axis=1)
    assert got.dtype == np.float32
    assert np.isnan(got[5]).all() and np.isnan(want[5]).all()      # no epsilon in the reference: 0/0
    m = np.ones(257, bool)
    m[5] = False
    np.testing.assert_allclose(got[m], want[m], atol=1e-6)
    np.testing.assert_allclose(hw.unit_length(a[3], axis=0), O.unit_length(a[3], axis=0), atol=1e-6)
    np.testing.assert_allclose(hw.unit_length(a[:, :5], axis=0)[:, 0], O.unit_length(a[:, :5], axis=0)[:, 0], atol=1e-6)


def test_unit_length_violations(hw, golden_c1):
    n_users, n_items, d, _ = [int(x) for x in golden_c1["shape"]]
    content, collab = synthetic_case(n_users, n_items, d, 100)
    table = O.blend_normalize(content, collab, 0.0)
    got = hw.unit_length_violations(table, axis=1)
    assert got[0] == 0 and got[2] == 0 and got[3] == 0
    p = table.copy()
    p[3] *= 1.01
    p[10] *= 0.9
    p[11] *= 1.0 + 5e-5
    got = hw.unit_length_violations(p, axis=1)
    ref = golden_c1["viol_perturbed"]
    assert (got[0], got[2], got[3]) == (int(ref[0]), int(ref[2]), int(ref[3]))
    assert abs(got[1] - ref[1]) < 1e-8


# ----------------------------------------------------------------------------- exact top-k
CASES = [
    # n,     d,   B,   k    (C1, C2, C3 item sides; C4/10; padded width; CUDA-core width; tiny)
    (1682, 64, 200, 10),
    (3706, 128, 256, 200),
    (27278, 256, 300, 100),
    (200000, 128, 64, 100),
    (1000000, 128, 7, 100),
    (50000, 100, 33, 50),
    (5000, 320, 20, 10),
    (300, 128, 1, 300),
    (129, 64, 3, 1),
]


@pytest.mark.parametrize("n,d,B,k", CASES)
def test_topk_exact_matches_oracle(hw, n, d, B, k):
    t_np, t = unit_table(n, d, 11)
    q_np, q = unit_table(B, d, 12)
    q_np[0] *= 0.37                                   # queries need not be unit length (:170 is not re-normalised)
    q = torch.from_numpy(q_np).cuda()
    idx, sc, s64 = hw.ops.TopKIndex(t).topk(q, k, "exact", want_f64=True)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    idx, sc, s64 = idx.cpu().numpy(), sc.cpu().numpy(), s64.cpu().numpy()
    assert O.compare_topk(idx, s64, ref_idx, ref_sc, tie_eps=1e-6) == 0
    assert (idx == ref_idx).mean() > 0.999
    np.testing.assert_allclose(s64, ref_sc, rtol=0, atol=1e-12)      # fp64 re-score
    np.testing.assert_allclose(sc, ref_sc, rtol=0, atol=1e-5)        # north_star: 1e-5 on fp32 scores
    assert np.all(np.diff(s64, axis=1) <= 0)


@pytest.mark.parametrize("n,d,B,k", [(200000, 128, 64, 100), (27278, 256, 100, 100)])
def test_topk_bf16_mode(hw, n, d, B, k):
    t_np, t = unit_table(n, d, 21)
    q_np, q = unit_table(B, d, 22)
    idx, sc = hw.ops.TopKIndex(t).topk(q, k, "bf16")
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    idx, sc = idx.cpu().numpy(), sc.cpu().numpy()
    recall = np.mean([len(set(a) & set(b)) / k for a, b in zip(idx, ref_idx)])
    assert recall >= 0.97, recall
    exact_of_returned = np.einsum("bkd,bd->bk", t_np[idx].astype(np.float64), q_np.astype(np.float64))
    assert np.max(np.abs(sc - exact_of_returned) / np.abs(exact_of_returned)) <= 2e-2     # north_star: 2e-2 relative
    # the bf16 answer is itself exact for the bf16-rounded operands
    tb = t.to(torch.bfloat16).float().cpu().numpy()
    qb = q.to(torch.bfloat16).float().cpu().numpy()
    bidx, bsc = O.exact_topk(tb, qb, k)
    assert O.compare_topk(idx, sc.astype(np.float64), bidx, bsc, tie_eps=2e-6) == 0


def test_duplicates_follow_row_order_and_overflow_retries(hw):
    """Exact duplicates are ordered by ascending row (the documented tie rule); more ties than the candidate
    capacity exercise the HWER_E_OVERFLOW -> larger cap retry."""
    n, d, k = 30000, 128, 100
    t_np, _ = unit_table(n, d, 31)
    t_np[5000:17000] = t_np[4999]                     # 12001 identical rows (> the automatic capacity of 8192 at B=300)
    t_np[20000:20003] = t_np[7]
    t = torch.from_numpy(t_np).cuda()
    q_np = np.concatenate([np.stack([t_np[4999], t_np[7], t_np[123]]), t_np[25000:25297]])
    q = torch.from_numpy(q_np).cuda()
    index = hw.ops.TopKIndex(t)
    idx, sc = index.topk(q, k)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    assert list(idx[0].cpu().numpy()) == list(range(4999, 4999 + k))
    assert list(idx[1, :4].cpu().numpy()) == [7, 20000, 20001, 20002]
    # the raw entry point reports the overflow instead of retrying
    from hwer_b200 import _native
    index.topk_async(q, k)
    rc, need = index.finish()
    assert rc == _native.HWER_E_OVERFLOW and need >= 12001
    # an explicit capacity that holds the tie group succeeds at once
    idx2, _, _ = index.topk_async(q, k, cap=16384)
    rc, need = index.finish()
    assert rc == _native.HWER_OK
    np.testing.assert_array_equal(idx2.cpu().numpy(), ref_idx)


@pytest.mark.parametrize("n,B", [(10000, 9), (9000, 300), (5000, 600), (4500, 1000)])
def test_catalogue_slightly_larger_than_dense_round(hw, n, B):
    """A table a little longer than the dense first round (8192 / 4096 rows) must not push that round past the
    candidate capacity: the raw entry point (no retry) has to succeed and be exact."""
    from hwer_b200 import _native
    t_np, t = unit_table(n, 128, 33)
    q_np, q = unit_table(B, 128, 34)
    index = hw.ops.TopKIndex(t)
    idx, sc, s64 = index.topk_async(q, 10, want_f64=True)
    rc, need = index.finish()
    assert rc == _native.HWER_OK, (rc, need)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, 10)
    assert O.compare_topk(idx.cpu().numpy(), s64.cpu().numpy(), ref_idx, ref_sc, tie_eps=1e-6) == 0


def test_sorted_clustered_catalogue(hw):
    """Rows sorted by cluster, queries aimed at the LAST cluster: the early rounds see other clusters only."""
    rs = np.random.RandomState(41)
    n_c, per, d, k = 64, 2000, 128, 100
    centers = O.unit_length(rs.standard_normal((n_c, d)).astype(np.float32), axis=1)
    t_np = np.repeat(centers, per, axis=0) + 0.05 * rs.standard_normal((n_c * per, d)).astype(np.float32)
    t_np = O.unit_length(t_np, axis=1)
    q_np = O.unit_length(centers[-3:] + 0.02 * rs.standard_normal((3, d)).astype(np.float32), axis=1)
    idx, sc, s64 = hw.ops.TopKIndex(torch.from_numpy(t_np).cuda()).topk(torch.from_numpy(q_np).cuda(), k, want_f64=True)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    assert O.compare_topk(idx.cpu().numpy(), s64.cpu().numpy(), ref_idx, ref_sc) == 0


def test_nan_rows_are_never_returned(hw):
    t_np, _ = unit_table(4000, 64, 51)
    t_np[17] = np.nan
    q_np, q = unit_table(4, 64, 52)
    idx, sc = hw.ops.TopKIndex(torch.from_numpy(t_np).cuda(), max_norm=1.0).topk(q, 10)
    assert 17 not in idx.cpu().numpy()
    keep = np.ones(4000, bool)
    keep[17] = False
    ref_idx, _ = O.exact_topk(t_np[keep], q_np, 10)
    remap = np.nonzero(keep)[0]
    np.testing.assert_array_equal(idx.cpu().numpy(), remap[ref_idx])


def test_k_larger_than_rows_raises_value_error(hw):
    _, t = unit_table(50, 64, 61)
    _, q = unit_table(2, 64, 62)
    with pytest.raises(ValueError):                   # sklearn's KDTree.query raises ValueError here
        hw.ops.TopKIndex(t).topk(q, 51)


# ----------------------------------------------------------------------------- the reference-facing API vs golden
@pytest.fixture(scope="module")
def c1_model(hw, golden_c1):
    n_users, n_items, d, k = [int(x) for x in golden_c1["shape"]]
    content, collab = synthetic_case(n_users, n_items, d, 100)
    users = [hw.Node("user", i) for i in range(n_users)]
    items = [hw.Node("item", i) for i in range(n_items)]
    edges = [hw.Edge(users[i % n_users], items[i % n_items], 1.0) for i in range(50)]
    base = hw.ContentRecommendation(None, {"user", "item"}, n_dims=d)
    table = O.blend_normalize(content, collab, 0.0)
    base.fit(users + items, edges, None, vectors=table)
    gcn = hw.GcnNCF(None, {"user", "item"}, n_dims=d, alpha=0.0)
    out = gcn.fit(users + items, edges, None, content_vectors=content, collaborative_vectors=collab)
    np.testing.assert_allclose(out, table, atol=1e-6)
    return dict(base=base, gcn=gcn, users=users, items=items, k=k, g=golden_c1)


def _ids(res):
    return [int(n.node_external_id) for n, s in res], [float(s) for n, s in res]


def test_batch_to_host_equals_the_plain_batch_call(c1_model):
    """find_closest_neighbours_batch_to_host (chunked, result copies overlapped with the next chunk's search) returns
    the rows, scores and order of the plain batched call, for Node anchors and for row tensors, any chunk size."""
    m, users = c1_model["base"], c1_model["users"]
    rows, sc = m.find_closest_neighbours_batch("item", users, k=10)
    for chunk in (17, 256, 100000):
        hr, hs = m.find_closest_neighbours_batch_to_host("item", users, k=10, chunk=chunk)
        assert hr.is_pinned() and not hr.is_cuda
        assert torch.equal(hr, rows.cpu()) and torch.equal(hs, sc.cpu())
    arows = m._rows_of(users)
    hr, hs = m.find_closest_neighbours_batch_to_host("item", arows, k=10, chunk=100)
    assert torch.equal(hr, rows.cpu()) and torch.equal(hs, sc.cpu())


def test_api_matches_reference_outputs(c1_model):
    m, g, k = c1_model["base"], c1_model["g"], c1_model["k"]
    users, items = c1_model["users"], c1_model["items"]
    for j, u in enumerate(g["user_anchors"]):
        idx, sc = _ids(m.find_items_for_user(users[int(u)], k=k))
        assert idx == list(g["items_for_user_idx"][j])
        np.testing.assert_allclose(sc, g["items_for_user_score"][j], atol=1e-6)
        assert _ids(m.find_closest_neighbours("item", users[int(u)], k=k))[0] == idx
    for j, i in enumerate(g["item_anchors"]):
        idx, sc = _ids(m.find_similar_items(items[int(i)], k=k))
        assert idx == list(g["similar_items_idx"][j]) and idx[0] == int(i)
        np.testing.assert_allclose(sc, g["similar_items_score"][j], atol=1e-6)
    for j, u in enumerate(g["user_anchors"][:8]):
        idx, _ = _ids(m.find_closest_neighbours("user", users[int(u)]))        # default k = 200
        assert idx == list(g["users_k200_idx"][j])
    n_items = len(items)
    for j, u in enumerate(g["user_anchors"][:16]):
        u = int(u)
        pos = [items[(u * 3 + t) % n_items] for t in range(3)]
        neg = [items[(u * 5 + t + 1) % n_items] for t in range(2)]
        idx, sc = _ids(m.find_closest_neighbours("item", users[u], positive=pos, negative=neg, k=k))
        assert idx == list(g["posneg_idx"][j])
        np.testing.assert_allclose(sc, g["posneg_score"][j], atol=1e-6)


def test_gcn_variant_and_multiknn_query_match_reference(c1_model, hw):
    m, g, k = c1_model["gcn"], c1_model["g"], c1_model["k"]
    users = c1_model["users"]
    for j, u in enumerate(g["user_anchors"][:32]):
        idx, sc = _ids(m.find_closest_neighbours("item", users[int(u)], k=k))
        assert idx == list(g["gcn_items_for_user_idx"][j])
        np.testing.assert_allclose(sc, g["gcn_items_for_user_score"][j], atol=1e-6)
    # MultiKNN.query: ascending Euclidean distance; (2 - dist) / 2 is the golden gcn score
    emb = m.get_average_embeddings([users[int(g["user_anchors"][0])]])
    res = m.knn.query(emb, "item", k=k)
    assert [int(n.node_external_id) for n, d_ in res] == list(g["gcn_items_for_user_idx"][0])
    np.testing.assert_allclose([(2 - d_) / 2 for n, d_ in res], g["gcn_items_for_user_score"][0], atol=1e-6)


def test_predict_matches_reference(c1_model, hw):
    m, g = c1_model["base"], c1_model["g"]
    nodes = c1_model["users"] + c1_model["items"]
    pairs = []
    for a, b in zip(g["pair_src"], g["pair_dst"]):
        pairs.append((nodes[a] if a >= 0 else hw.Node("user", "ghost%d" % len(pairs)),
                      nodes[b] if b >= 0 else hw.Node("item", "ghost%d" % len(pairs))))
    got = m.predict(pairs)
    np.testing.assert_allclose(got, g["pair_pred"], atol=1e-6)
    assert not np.isnan(got).any()


def test_batch_equals_per_anchor_and_errors(c1_model, hw):
    m, users = c1_model["base"], c1_model["users"]
    anchors = [users[i] for i in (3, 17, 400, 942)]
    rows, scores = m.find_closest_neighbours_batch("item", anchors, k=25)
    for a, r, s in zip(anchors, m.rows_to_nodes(rows, scores), scores.cpu().numpy()):
        one = m.find_closest_neighbours("item", a, k=25)
        assert [n for n, _ in r] == [n for n, _ in one]
        np.testing.assert_allclose([x for _, x in one], s, atol=1e-7)
    gm = c1_model["gcn"]
    rows, scores = gm.find_closest_neighbours_batch("item", anchors, k=25)
    one = gm.find_closest_neighbours("item", anchors[1], k=25)
    assert [n for n, _ in gm.rows_to_nodes(rows, scores)[1]] == [n for n, _ in one]
    np.testing.assert_allclose([x for _, x in one], scores[1].cpu().numpy(), atol=1e-6)
    with pytest.raises(hw.NodeNotFoundException):
        m.find_closest_neighbours("item", hw.Node("user", "nobody"))
    with pytest.raises(AssertionError):
        m.find_closest_neighbours("genre", users[0])
    with pytest.raises(ValueError):
        m.find_closest_neighbours("item", users[0], k=5000)
    with pytest.raises(AssertionError):                               # unit-norm violation (:106-107)
        bad = hw.ContentRecommendation(None, {"user", "item"}, n_dims=64)
        bad.add_nodes(users[:10])
        bad.__build_knn__(np.full((10, 64), 0.5, dtype=np.float32))


def test_interleaved_node_types_are_gathered(hw):
    """Types interleaved in the node list (the general case of MultiKNN.__init__, :68-74)."""
    n, d = 600, 64
    t_np, _ = unit_table(n, d, 71)
    nodes = [hw.Node("item" if i % 3 else "user", i) for i in range(n)]
    edges = [hw.Edge(nodes[0], nodes[1], 1.0)]
    m = hw.ContentRecommendation(None, {"user", "item"}, n_dims=d)
    m.fit(nodes, edges, None, vectors=t_np)
    item_rows = np.array([i for i in range(n) if i % 3])
    ref_idx, ref_sc = O.exact_topk(t_np[item_rows], t_np[[0]], 10)
    got = m.find_closest_neighbours("item", nodes[0], k=10)
    assert [int(x.node_external_id) for x, _ in got] == list(item_rows[ref_idx[0]])


# ----------------------------------------------------------------------------- evaluation kernel
def test_extraction_efficiency_matches_reference_metrics(hw, golden_eval):
    nu, ni, dd = [int(x) for x in golden_eval["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_eval["seeds"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(golden_eval["seeds"][1]))
    train = [hw.Edge(users[u], items[i], w) for u, i, w in tr]
    val = [hw.Edge(users[u], items[i], w) for u, i, w in vl]
    m = hw.GcnNCF(None, {"user", "item"}, n_dims=dd)
    m.fit(users + items, train, None, collaborative_vectors=collab)
    base = hw.ContentRecommendation(None, {"user", "item"}, n_dims=dd)
    base.fit(users + items, train, None, vectors=O.unit_length(collab, axis=1))
    res = hw.validation.extraction_efficiency(base, train, val, hw.validation.model_get_topk, "item")
    ref = dict(zip([str(k) for k in golden_eval["metric_keys"]], golden_eval["metric_values"]))
    for key in ("recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"):
        assert abs(res["metrics"][key] - ref[key]) < 1e-9, (key, res["metrics"][key], ref[key])
    assert 0.0 <= res["metrics"]["ncf_hr"] <= 1.0 and res["metrics"]["retrieval_time"] > 0
    # all cutoffs + graded ndcg + mrr against the oracle's restatement of validation.py
    ou = [O.Node("user", i) for i in range(nu)]
    oi = [O.Node("item", i) for i in range(ni)]
    om = O.OracleRecommender({"user", "item"}, n_dims=dd)
    om.add_nodes(ou + oi)
    om.build_knn(O.unit_length(collab, axis=1))
    otr = [(ou[u], oi[i], w) for u, i, w in tr]
    ovl = [(ou[u], oi[i], w) for u, i, w in vl]
    all_users = list(set([u for u, i, r in otr] + [u for u, i, r in ovl]))
    want = O.extraction_metrics(O.model_get_topk_knn(om, all_users, "item"), otr, ovl, "item")
    for key, v in want.items():
        got = res["all_metrics"][key] if key != "diversity" else res["metrics"]["diversity"]
        assert abs(got - v) < 1e-9, (key, got, v)
    # the dict-returning hook keeps the reference's return type
    d = hw.validation.model_get_topk(base, [users[0], users[5]], "item")
    sample = dict(zip(golden_eval["sample_users"], golden_eval["sample_preds"]))
    train_items = set(i for u, i, w in tr if u == 5)
    mine = [int(n.node_external_id) for n, s in d[users[5]] if int(n.node_external_id) not in train_items][:100]
    assert mine == [int(x) for x in sample[5] if x >= 0]


# ----------------------------------------------------------------------------- multi-GPU pieces on one GPU
def test_merge_and_shard_equivalence(hw):
    n, d, B, k = 100000, 128, 40, 100
    t_np, t = unit_table(n, d, 81)
    t_np[60000:60004] = t_np[100]                      # ties across the shard boundary
    t = torch.from_numpy(t_np).cuda()
    _, q = unit_table(B, d, 82)
    whole_idx, whole_sc, whole_s64 = hw.ops.TopKIndex(t).topk(q, k, want_f64=True)
    parts = []
    for g in range(3):
        b, e = hw.sharded.partition(n, 3, g)
        sh = hw.sharded.ShardedTopK(t[b:e].contiguous(), b)
        parts.append(sh.local_topk(q, k))
    gi = torch.stack([p[0] for p in parts])
    gs = torch.stack([p[1] for p in parts])
    idx, sc, s64 = hw.ops.merge_topk(gs.contiguous(), gi.contiguous(), want_f64=True)
    assert torch.equal(idx, whole_idx) and torch.equal(s64, whole_s64)      # identical for any shard count
    # merge kernel vs numpy lexsort
    gs_n, gi_n = gs.cpu().numpy(), gi.cpu().numpy()
    for r in range(B):
        s, i = gs_n[:, r].reshape(-1), gi_n[:, r].reshape(-1)
        order = np.lexsort((i, -s))[:k]
        np.testing.assert_array_equal(idx[r].cpu().numpy(), i[order])


class _Exchange:
    """Bare exchange handle for topk_sharded_async (the product's PeerExchange also does the IPC rendezvous)."""
    def __init__(self, h):
        self._h = h


# (one GPU plays every rank here, so batches stay small: a rank's waiting kernels must leave SMs free for the kernels
#  of the ranks it waits for -- with one GPU per rank, as deployed, that coupling does not exist)
@pytest.mark.parametrize("world,B,k,share", [(2, 50, 100, 1), (3, 7, 10, 1), (1, 33, 20, 1), (3, 40, 100, 1),
                                             (2, 50, 100, 0), (3, 9, 10, 0), (2, 12, 1000, 1)])
def test_peer_exchange_protocol_on_one_gpu(hw, world, B, k, share):
    """hwer_topk_sharded with `world` ranks living in ONE process: one exchange buffer, index (a row shard) and
    stream per rank.  Exercises the real protocol -- per-round threshold sharing (share=1), final kernel storing
    into the owner's buffer, release/acquire flags, owner merge delivering to every rank, collect -- and must
    reproduce the single-index answer bit for bit on every rank, twice (buffer reuse across epochs)."""
    import ctypes
    from hwer_b200 import _native as N
    lib = N.lib()
    n, d = 30011, 128                                    # shards of unequal size: a common round schedule matters
    t_np, t = unit_table(n, d, 91)
    t_np[20000:20003] = t_np[5]                          # ties across shard boundaries
    t = torch.from_numpy(t_np).cuda()
    whole = hw.ops.TopKIndex(t, max_norm=1.0001)
    b_cap = max(256, B)
    nbytes = lib.hwer_exchange_bytes(world, b_cap, k)
    assert nbytes > 0
    bases = (ctypes.c_void_p * world)()
    handle = (ctypes.c_ubyte * N.IPC_HANDLE_BYTES)()
    for r in range(world):
        ptr = ctypes.c_void_p()
        N.check(lib.hwer_peer_alloc(nbytes, ctypes.byref(ptr), handle))
        bases[r] = ptr
    ex, shards, streams = [], [], []
    for r in range(world):
        h = ctypes.c_void_p()
        N.check(lib.hwer_exchange_create(ctypes.byref(h), world, r, b_cap, k, bases, 0))
        rows_max = max(hw.sharded.partition(n, world, g)[1] - hw.sharded.partition(n, world, g)[0] for g in range(world))
        N.check(lib.hwer_exchange_configure(h, rows_max, share))
        ex.append(_Exchange(h))
        b, e = hw.sharded.partition(n, world, r)
        shards.append((hw.ops.TopKIndex(t[b:e].contiguous(), max_norm=1.0001), b))
        shards[-1][0].topk(t[:B].contiguous(), k)        # sizes the workspace now (growing it synchronises the device)
        streams.append(torch.cuda.Stream())
    try:
        for seed in (92, 93):
            _, q = unit_table(B, d, seed)
            ref_idx, ref_sc, ref_s64 = whole.topk(q, k, want_f64=True)
            torch.cuda.synchronize()
            outs = [(torch.empty((B, k), dtype=torch.int64, device="cuda"),
                     torch.empty((B, k), dtype=torch.float32, device="cuda"),
                     torch.empty((B, k), dtype=torch.float64, device="cuda")) for _ in range(world)]
            torch.cuda.synchronize()
            # one process plays every rank, so the phases are enqueued rank by rank, phase by phase: a rank's merge
            # spins on the GPU until the other ranks' searches (already enqueued on their streams) have published
            for phase in (N.PHASE_SEARCH, N.PHASE_MERGE, N.PHASE_COLLECT):
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        shards[r][0].topk_sharded_async(ex[r], q, k, idx_offset=shards[r][1], phases=phase, out=outs[r])
            torch.cuda.synchronize()
            for r in range(world):
                N.check(lib.hwer_exchange_error(ex[r]._h, None))
                assert torch.equal(outs[r][0], ref_idx), "rank %d rows differ" % r
                assert torch.equal(outs[r][2], ref_s64) and torch.equal(outs[r][1], ref_sc)
            # HWER_PHASE_OWNED: every rank keeps only the queries it merged, packed from row 0 of its outputs
            per = (B + world - 1) // world
            for o in outs:
                o[0].fill_(-7)
            for phase in (N.PHASE_SEARCH, N.PHASE_MERGE | N.PHASE_OWNED, N.PHASE_COLLECT | N.PHASE_OWNED):
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        shards[r][0].topk_sharded_async(ex[r], q, k, idx_offset=shards[r][1], phases=phase, out=outs[r])
            torch.cuda.synchronize()
            for r in range(world):
                N.check(lib.hwer_exchange_error(ex[r]._h, None))
                lo, hi = min(r * per, B), min((r + 1) * per, B)
                assert torch.equal(outs[r][0][:hi - lo], ref_idx[lo:hi]), "rank %d owned rows differ" % r
                assert torch.equal(outs[r][2][:hi - lo], ref_s64[lo:hi]) and torch.equal(outs[r][1][:hi - lo], ref_sc[lo:hi])
                assert bool((outs[r][0][hi - lo:] == -7).all()), "rank %d wrote past its owned rows" % r
    finally:
        torch.cuda.synchronize()
        for r in range(world):
            lib.hwer_exchange_destroy(ex[r]._h)
            lib.hwer_peer_free(bases[r])


def _p2p_worker(rank, world, port, out_dir):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import hwer_b200 as hwm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n, d, B, k = 200000, 128, 300, 100
    g = torch.Generator(device="cuda").manual_seed(7)
    table = hwm.ops.unit_length(torch.randn((n, d), generator=g, device="cuda"))
    q = hwm.ops.unit_length(torch.randn((B, d), generator=g, device="cuda"))
    b, e = hwm.sharded.partition(n, world, rank)
    res = {}
    for ex in ("p2p", "nccl"):
        sh = hwm.sharded.ShardedTopK(table[b:e].contiguous(), b, exchange=ex)
        for _ in range(2):
            idx, sc = sh.topk(q, k)
        res[ex] = (idx.cpu(), sc.cpu())
        # owned=True: each rank gets exactly the rows of the queries it merged
        lo, hi = sh.owner_range(B)
        oidx, osc = sh.topk(q, k, owned=True)
        assert oidx.shape[0] == hi - lo and torch.equal(oidx.cpu(), res[ex][0][lo:hi]) and torch.equal(osc.cpu(), res[ex][1][lo:hi])
        sh.close()
    whole_idx, whole_sc = hwm.ops.TopKIndex(table).topk(q, k)
    assert torch.equal(res["p2p"][0], whole_idx.cpu()) and torch.equal(res["nccl"][0], whole_idx.cpu())
    assert torch.equal(res["p2p"][1], res["nccl"][1])
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_peer_exchange_equals_nccl_and_single_gpu(tmp_path):
    import os
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 1500)
    mp.spawn(_p2p_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))


# ----------------------------------------------------------------------------- query composition
@pytest.mark.parametrize("d", [64, 100, 256])
def test_compose_queries_matches_oracle(hw, d):
    """hwer_compose_queries against the oracle's restatement of recommendation_base.py:164-170 (itself pinned to
    the reference's positive/negative outputs in tests/golden/reference_c1.npz)."""
    rs = np.random.RandomState(71)
    n, B = 500, 40
    t_np, t = unit_table(n, d, 72)
    nodes = [O.Node("item", i) for i in range(n)]
    rec = O.OracleRecommender({"item"}, n_dims=d)
    rec.add_nodes(nodes)
    rec.build_knn(t_np)
    ghost = O.Node("item", "ghost")
    anchors, pos, neg = [], [], []
    for b in range(B):
        anchors.append(int(rs.randint(n)))
        pos.append([int(x) for x in rs.randint(0, n, rs.randint(0, 4))] if b % 3 else [])
        neg.append([int(x) for x in rs.randint(0, n, rs.randint(0, 3))] if b % 2 else [])
    pos[5] = pos[5] + [-1]                              # a node that was never trained on
    want = np.stack([rec.query_embedding(nodes[a], [nodes[i] if i >= 0 else ghost for i in p] or None,
                                         [nodes[i] for i in ng] or None) for a, p, ng in zip(anchors, pos, neg)])

    def csr(lists):
        ptr = np.cumsum([0] + [len(x) for x in lists]).astype(np.int64)
        rows = np.array([i for x in lists for i in x], dtype=np.int64)
        return torch.from_numpy(ptr).cuda(), torch.from_numpy(rows).cuda()

    got = hw.ops.compose_queries(t, torch.tensor(anchors, dtype=torch.int64, device="cuda"), csr(pos), csr(neg))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-6)
    only_anchor = hw.ops.compose_queries(t, torch.tensor(anchors, dtype=torch.int64, device="cuda"))
    np.testing.assert_allclose(only_anchor.cpu().numpy(), t_np[anchors], rtol=0, atol=2e-6)


def test_batch_posneg_equals_per_anchor(c1_model, hw):
    model, users, items, g = c1_model["base"], c1_model["users"], c1_model["items"], c1_model["g"]
    anchors = [users[int(u)] for u in g["user_anchors"][:16]]
    n_items = len(items)
    pos = [[items[(int(u) * 3 + j) % n_items] for j in range(3)] for u in g["user_anchors"][:16]]
    neg = [[items[(int(u) * 5 + j + 1) % n_items] for j in range(2)] for u in g["user_anchors"][:16]]
    k = int(g["shape"][3])
    rows, sc = model.find_closest_neighbours_batch("item", anchors, k=k, positive=pos, negative=neg)
    got = rows.cpu().numpy() - len(users)
    np.testing.assert_array_equal(got, g["posneg_idx"])                 # the reference's own outputs
    np.testing.assert_allclose(sc.cpu().numpy(), g["posneg_score"], rtol=0, atol=1e-5)


def test_table_file_to_device_and_serve(hw, tmp_path):
    """save_table -> load_table(device='cuda', one shard's rows) -> index: same answers as the in-memory table."""
    import os
    t_np, t = unit_table(20000, 128, 75)
    _, q = unit_table(9, 128, 76)
    p = os.path.join(str(tmp_path), "items.hwer")
    hw.table_io.save_table(p, t, node_types={"item": (0, 20000)})
    full, hdr = hw.table_io.load_table(p, device="cuda", chunk_rows=3000)
    assert torch.equal(full, t) and hdr["unit_norm"]
    want = hw.ops.TopKIndex(t).topk(q, 10)
    got = hw.ops.TopKIndex(full).topk(q, 10)
    assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])
    half, _ = hw.table_io.load_table(p, device="cuda", rows=(10000, 20000), chunk_rows=4096)
    assert torch.equal(half, t[10000:])


# ----------------------------------------------------------------------------- NCF re-rank
@pytest.fixture(scope="module")
def golden_ncf():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "reference_ncf.npz"))


@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_ncf_score_matches_reference(hw, golden_ncf, depth):
    """hwer_ncf_score against the reference NCF module's own outputs (tests/golden/reference_ncf.npz): 1e-5."""
    g = golden_ncf
    out = hw.ops.ncf_score(torch.from_numpy(g["h"]).cuda(), torch.from_numpy(g["params_d%d" % depth]).cuda(),
                           torch.from_numpy(g["src"]).cuda(), torch.from_numpy(g["dst"]).cuda(), depth)
    np.testing.assert_allclose(out.cpu().numpy(), g["forward_d%d" % depth], rtol=0, atol=1e-5)


@pytest.mark.parametrize("F,depth,P", [(128, 3, 70001), (64, 2, 333), (36, 4, 1000), (256, 3, 4097)])
def test_ncf_score_matches_oracle_at_serving_widths(hw, F, depth, P):
    rs = np.random.RandomState(61)
    n = 5000
    h = (rs.standard_normal((n + 1, F)) * 0.3).astype(np.float32)
    params = np.concatenate([np.concatenate([(rs.standard_normal((o, i)) / np.sqrt(i)).reshape(-1),
                                             rs.standard_normal(o) * 0.1]) for i, o in O.ncf_layer_dims(F, depth)])
    params = params.astype(np.float32)
    src = rs.randint(-1, n + 3, P).astype(np.int64)             # includes out-of-range rows (-> padding row 0)
    dst = rs.randint(0, n + 1, P).astype(np.int64)
    out = hw.ops.ncf_score(torch.from_numpy(h).cuda(), torch.from_numpy(params).cuda(), torch.from_numpy(src).cuda(),
                           torch.from_numpy(dst).cuda(), depth).cpu().numpy()
    src_c = np.where((src < 0) | (src > n), 0, src)
    ref = O.ncf_forward(h, src_c, dst, params, depth)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-5)


def test_gcn_ncf_rerank_matches_reference(hw, golden_ncf):
    """GcnNCF with the NCF branch enabled: predict (incl. unknown nodes) and find_closest_neighbours re-ranked by
    the NCF, against the reference's own outputs (gcn_ncf.py:336-361, 363-387)."""
    g = golden_ncf
    n_users, n_items, F, k = [int(x) for x in g["shape"]]
    users = [hw.Node("user", i) for i in range(n_users)]
    items = [hw.Node("item", i) for i in range(n_items)]
    edges = [hw.Edge(users[i % n_users], items[i % n_items], 1.0) for i in range(100)]
    model = hw.GcnNCF(None, {"user", "item"}, n_dims=F)
    model.fit(users + items, edges, None, collaborative_vectors=g["table"])
    model.set_ncf(g["h"], g["params_d3"], 3)
    inv = model.nodes_to_idx.inverse
    pairs = []
    for a, b in zip(g["predict_src"], g["predict_dst"]):
        pairs.append((inv[int(a)] if a >= 0 else hw.Node("user", "ghost"), inv[int(b)] if b >= 0 else hw.Node("item", "ghost")))
    pred = model.predict(pairs)
    assert isinstance(pred, list)
    np.testing.assert_allclose(np.asarray(pred, dtype=np.float64), g["predict"], rtol=0, atol=1e-5)
    for j, u in enumerate(g["anchors"]):
        res = model.find_closest_neighbours("item", users[int(u)], k=k)
        assert [int(n.node_external_id) for n, s in res] == list(g["fcn_idx"][j])
        np.testing.assert_allclose([s for n, s in res], g["fcn_score"][j], rtol=0, atol=1e-5)
    rows, sc = model.find_closest_neighbours_batch("item", [users[int(u)] for u in g["anchors"]], k=k)
    np.testing.assert_array_equal(rows.cpu().numpy() - n_users, g["fcn_idx"])
    np.testing.assert_allclose(sc.cpu().numpy(), g["fcn_score"], rtol=0, atol=1e-5)


def test_scaled_and_biased_filter_blocks_agree(hw, monkeypatch):
    """The filter rounds test score / thr >= 1 on query blocks whose thresholds are all positive (queries staged as
    q / thr: no extra MMA) and fall back to the bias MMA (score - thr >= 0) otherwise.  Both must return the oracle's
    rows: a clustered catalogue where some queries only see negative scores, huge / tiny query norms, mixed blocks."""
    n, d, k = 300000, 128, 50
    rs = np.random.RandomState(55)
    centre = rs.standard_normal(d).astype(np.float32)
    t_np = O.unit_length((centre[None, :] + 0.7 * rs.standard_normal((n, d))).astype(np.float32), axis=1)
    q_np = O.unit_length(rs.standard_normal((300, d)).astype(np.float32), axis=1)
    q_np[:40] = -O.unit_length((centre[None, :] + 0.5 * rs.standard_normal((40, d))).astype(np.float32), axis=1)  # all scores < 0
    q_np[40:60] = O.unit_length((centre[None, :] + 0.5 * rs.standard_normal((20, d))).astype(np.float32), axis=1)
    q_np[60:70] *= 1e6                                                    # any norm is allowed
    q_np[70:80] *= 1e-6
    q_np[299] = 0.0                                                       # zero query: every score is 0
    t, q = torch.from_numpy(t_np).cuda(), torch.from_numpy(q_np).cuda()
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    answers = []
    for disable in ("0", "1"):
        monkeypatch.setenv("HWER_DISABLE", disable)                       # read when the index is created
        index = hw.ops.TopKIndex(t)
        for sl in (slice(0, 300), slice(0, 40), slice(40, 300), slice(100, 101)):
            idx, sc, s64 = index.topk(q[sl].contiguous(), k, want_f64=True)
            assert O.compare_topk(idx.cpu().numpy(), s64.cpu().numpy(), ref_idx[sl], ref_sc[sl]) == 0, (disable, sl)
        answers.append(index.topk(q, k, want_f64=True))
    assert torch.equal(answers[0][0], answers[1][0]) and torch.equal(answers[0][2], answers[1][2])


def test_more_ties_than_the_selector_holds_falls_back_to_exhaustive_search(hw):
    """more than 16384 rows inside the bf16 margin of the k-th score (duplicated cold-start embeddings)
    used to raise; the reference's KDTree always answers.  Now the marked queries are answered exhaustively --
    same rows (ascending among exact duplicates), same fp64 scores -- while the other queries keep their result."""
    n, d, k = 60000, 128, 100
    t_np, _ = unit_table(n, d, 81)
    q_np, _ = unit_table(6, d, 82)
    dup = np.arange(5000, 5000 + 20000)
    t_np[dup] = q_np[2]                                       # 20,000 copies of query 2
    t = torch.from_numpy(t_np).cuda()
    q = torch.from_numpy(q_np).cuda()
    index = hw.ops.TopKIndex(t)
    idx, sc, s64 = index.topk(q, k, want_f64=True)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    assert idx[2].cpu().tolist() == list(range(5000, 5000 + k))
    np.testing.assert_allclose(s64.cpu().numpy(), ref_sc, rtol=0, atol=1e-6)
    # the exhaustive path alone: bit-identical to the filtered path on ordinary queries
    e_idx, e_sc, e_s64 = index.topk_exhaustive(q[[0, 1, 3]].contiguous(), k)
    assert torch.equal(e_idx, idx[[0, 1, 3]]) and torch.equal(e_s64, s64[[0, 1, 3]])
    t2_np, t2 = unit_table(3000, 100, 83)                     # scalar-dot width (d % 128 != 0), k close to n
    q2_np, q2 = unit_table(4, 100, 84)
    i2 = hw.ops.TopKIndex(t2)
    a = i2.topk(q2, 1000, want_f64=True)
    b = i2.topk_exhaustive(q2, 1000, idx_offset=7)
    assert torch.equal(a[0] + 7, b[0]) and torch.equal(a[2], b[2])


# ----------------------------------------------------------------------------- round 2: reference_r2.npz
def test_gcn_posneg_single_and_batch_match_reference(c1_model, golden_r2, hw):
    """GcnNCF.find_closest_neighbours with positive / negative lists (hwer/gcn_ncf.py:363-383): (2 - dist) / 2 with
    dist measured to the COMPOSED embedding -- per anchor and batched, against the reference's own outputs."""
    m, g, k = c1_model["gcn"], c1_model["g"], c1_model["k"]
    users, items = c1_model["users"], c1_model["items"]
    anchors = [users[int(u)] for u in g["user_anchors"][:16]]
    pn = [posneg_lists(int(u), len(items)) for u in g["user_anchors"][:16]]
    pos = [[items[t] for t in p] for p, n in pn]
    neg = [[items[t] for t in n] for p, n in pn]
    for j, a in enumerate(anchors):
        idx, sc = _ids(m.find_closest_neighbours("item", a, positive=pos[j], negative=neg[j], k=k))
        assert idx == list(golden_r2["gcn_posneg_idx"][j])
        np.testing.assert_allclose(sc, golden_r2["gcn_posneg_score"][j], rtol=0, atol=1e-6)
    rows, sc = m.find_closest_neighbours_batch("item", anchors, k=k, positive=pos, negative=neg)
    np.testing.assert_array_equal(rows.cpu().numpy() - len(users), golden_r2["gcn_posneg_idx"])
    np.testing.assert_allclose(sc.cpu().numpy(), golden_r2["gcn_posneg_score"], rtol=0, atol=1e-6)
    # positives only / negatives only, default k = 200, batched with ragged lists (None entries)
    one_anchor, one_pos, one_neg = [], [], []
    for u in g["user_anchors"][16:20]:
        p, n = posneg_lists(int(u), len(items))
        one_anchor += [users[int(u)], users[int(u)]]
        one_pos += [[items[t] for t in p], None]
        one_neg += [None, [items[t] for t in n]]
    rows, sc = m.find_closest_neighbours_batch("item", one_anchor, positive=one_pos, negative=one_neg)
    np.testing.assert_array_equal(rows.cpu().numpy() - len(users), golden_r2["gcn_one_sided_idx"])
    np.testing.assert_allclose(sc.cpu().numpy(), golden_r2["gcn_one_sided_score"], rtol=0, atol=1e-6)


def test_knn_query_and_embeddings_match_reference(c1_model, golden_r2, hw):
    m, users, items = c1_model["base"], c1_model["users"], c1_model["items"]
    u0 = int(golden_r2["knn_query_user"][0])
    p, n = posneg_lists(u0, len(items))
    emb = m._query_embedding(users[u0], [items[t] for t in p], [items[t] for t in n])
    res = m.knn.query(emb, "item", k=25)                                    # MultiKNN.query, :78-83
    assert [int(x.node_external_id) for x, _ in res] == list(golden_r2["knn_query_idx"])
    np.testing.assert_allclose([d for _, d in res], golden_r2["knn_query_dist"], rtol=0, atol=1e-6)
    assert all(a <= b for (_, a), (_, b) in zip(res, res[1:]))
    probe = [users[3], items[7], hw.Node("user", "ghost"), items[1681], hw.Node("item", "ghost2")]
    np.testing.assert_array_equal(m.get_embeddings(probe), golden_r2["emb_rows"])       # :146-151, exact copies
    np.testing.assert_allclose(m.get_average_embeddings([items[1], items[2], items[3]]), golden_r2["avg_a"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(m.get_average_embeddings([users[5], hw.Node("item", "ghost3")]), golden_r2["avg_b"],
                               rtol=0, atol=1e-6)


def _eval_models(hw, golden_eval):
    nu, ni, dd = [int(x) for x in golden_eval["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_eval["seeds"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(golden_eval["seeds"][1]))
    train = [hw.Edge(users[u], items[i], w) for u, i, w in tr]
    val = [hw.Edge(users[u], items[i], w) for u, i, w in vl]
    base = hw.ContentRecommendation(None, {"user", "item"}, n_dims=dd)
    base.fit(users + items, train, None, vectors=O.unit_length(collab, axis=1))
    return base, users, items, train, val, collab, tr, vl


def test_ncf_eval_matches_reference(hw, golden_eval, golden_r2):
    """validation.ncf_eval (hwer/validation.py:68-97) under the reference's seed: same negatives, device pair
    scores, device rank / HR@10 / NDCG@10 -- equal to the reference's own numbers, and rank by rank to the oracle."""
    import random
    base, users, items, train, val, collab, tr, vl = _eval_models(hw, golden_eval)
    all_items = [x for x in set([i for u, i, w in val] + [i for u, i, w in train]) if x.node_type == "item"]
    random.seed(int(golden_r2["ncf_eval_seed"][0]))
    got = hw.validation.ncf_eval(base, train, val, all_items)
    np.testing.assert_allclose([got["ncf_hr"], got["ncf_ndcg"]], golden_r2["ncf_eval"], rtol=0, atol=1e-9)
    # device ranks against the oracle's per-user ranks on identical scores
    rs = np.random.RandomState(3)
    s_np = rs.rand(500, 101).astype(np.float32)
    s_np[::7, 5] = s_np[::7, 0]                       # ties with the positive: the positive stays ahead
    out, rank = hw.ops.hit_rank_metrics(torch.from_numpy(s_np).cuda(), topn=10, want_rank=True)
    want_rank = (s_np[:, 1:] > s_np[:, :1]).sum(1)
    np.testing.assert_array_equal(rank.cpu().numpy(), want_rank)
    hit = want_rank < 10
    np.testing.assert_allclose(out.cpu().numpy(), [hit.mean(), np.where(hit, 1 / np.log2(want_rank + 2.0) / (1 + 1e-8), 0).mean()],
                               rtol=0, atol=1e-12)


def test_prepare_for_knn_pca_branch_matches_reference(hw, golden_r2):
    """GcnNCF.prepare_for_knn reduces a table wider than n_dims with PCA (hwer/gcn_ncf.py:449-452) before the unit
    normalisation; the reference's own output on the same 400 x 96 table."""
    rs = np.random.RandomState(700)
    wide = (rs.standard_normal((400, 96)) * np.linspace(3.0, 0.2, 96)[None, :]).astype(np.float32)
    m = hw.GcnNCF(None, {"user", "item"}, n_dims=32)
    got = m.prepare_for_knn(None, wide)
    assert isinstance(got, np.ndarray) and got.shape == (400, 32)
    # sklearn decomposes the fp32 table in fp32: axes with close eigenvalues rotate by ~1e-4 against the float64
    # decomposition, while the Gram matrix -- all that retrieval sees -- agrees to 1e-5
    np.testing.assert_allclose(got, golden_r2["pca_table"], rtol=0, atol=3e-4)
    np.testing.assert_allclose(got @ got.T, golden_r2["pca_table"] @ golden_r2["pca_table"].T, rtol=0, atol=2e-5)
    with pytest.raises(ValueError):
        m.prepare_for_knn(None, wide[:, :16])


def test_get_topk_hook_is_honoured(hw, golden_eval):
    """extraction_efficiency(model, train, val, get_topk, node_type) calls the hook it is given
    (hwer/validation.py:100,111) and evaluates what it returns."""
    base, users, items, train, val, collab, tr, vl = _eval_models(hw, golden_eval)
    calls = []

    def reference_style_hook(model, anchors, node_type):           # hwer/validation.py:30-35
        calls.append(len(anchors))
        return {u: model.find_closest_neighbours(node_type, u) for u in anchors}

    ref = dict(zip([str(k) for k in golden_eval["metric_keys"]], golden_eval["metric_values"]))
    res = hw.validation.extraction_efficiency(base, train, val, reference_style_hook, "item")
    assert calls and calls[0] == len(set(u for u, i, w in tr) | set(u for u, i, w in vl))
    for key in ("recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"):
        assert abs(res["metrics"][key] - ref[key]) < 1e-9

    def empty_hook(model, anchors, node_type):
        return {}
    res = hw.validation.extraction_efficiency(base, train, val, empty_hook, "item")
    assert res["metrics"]["recall@100"] == 0.0 and res["metrics"]["diversity"] == 0.0


@pytest.mark.parametrize("B,k,d", [(7, 10, 64), (3, 200, 100), (130, 100, 128), (2, 1000, 32), (1, 1, 8)])
def test_rerank_conventions(hw, B, k, d):
    """hwer_rerank against numpy: the four score conventions, the stable order, missing rows, a row map."""
    n = max(2 * k, 500)
    t_np, t = unit_table(n, d, 91)
    rs = np.random.RandomState(92)
    rows = np.stack([rs.choice(n, k, replace=False) for _ in range(B)]).astype(np.int64)
    if k > 4:
        rows[0, k - 2:] = -1                                   # missing results stay last
        rows[-1, 1] = rows[-1, 0]                              # a duplicate: equal scores keep the input order
    anchors = rs.randint(0, n, B).astype(np.int64)
    q_np = (t_np[anchors] + 0.3 * rs.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    rows_t, anchors_t, q_t = torch.from_numpy(rows).cuda(), torch.from_numpy(anchors).cuda(), torch.from_numpy(q_np).cuda()
    safe = np.where(rows >= 0, rows, 0)

    def check(conv, scores, descending, **kw):
        got_rows, got_sc = hw.ops.rerank(t, rows_t, conv, **kw)
        got_rows, got_sc = got_rows.cpu().numpy(), got_sc.cpu().numpy()
        for b in range(B):
            valid = np.flatnonzero(rows[b] >= 0)
            key = -scores[b, valid] if descending else scores[b, valid]
            order = valid[np.argsort(key, kind="stable")]
            np.testing.assert_array_equal(got_rows[b, :len(order)], rows[b, order])
            np.testing.assert_allclose(got_sc[b, :len(order)], scores[b, order], rtol=0, atol=1e-6)
            assert (got_rows[b, len(order):] == -1).all()
            assert np.isinf(got_sc[b, len(order):]).all()

    pair = ((t_np[anchors][:, None, :] * t_np[safe]).sum(2) + 1) / 2
    pair_dev = hw.ops.pair_score(t, anchors_t[:, None].expand(B, k).reshape(-1).contiguous(),
                                 torch.from_numpy(safe).cuda().reshape(-1)).reshape(B, k).cpu().numpy()
    check("pair", pair_dev.astype(np.float64), True, anchor_rows=anchors_t)     # bit-equal to predict()'s kernel
    np.testing.assert_allclose(pair_dev, pair, atol=1e-6)
    dist = np.sqrt(((t_np[safe].astype(np.float64) - q_np[:, None, :].astype(np.float64)) ** 2).sum(2))
    check("dist", (2 - dist) / 2, True, queries=q_t)
    check("euclid", dist, False, queries=q_t)
    given = rs.rand(B, k).astype(np.float32)
    given[:, k // 2:] = given[:, :k - k // 2]                   # many exact ties
    check("given", given.astype(np.float64), True, given=torch.from_numpy(given).cuda())
    # rows local to a gathered index: out rows are the mapped (global) ones
    perm = rs.permutation(n).astype(np.int64)
    got_rows, got_sc = hw.ops.rerank(t, rows_t, "dist", queries=q_t, row_map=torch.from_numpy(perm).cuda())
    dist_m = np.sqrt(((t_np[perm[safe]].astype(np.float64) - q_np[:, None, :].astype(np.float64)) ** 2).sum(2))
    for b in range(B):
        valid = np.flatnonzero(rows[b] >= 0)
        order = valid[np.argsort(-((2 - dist_m[b, valid]) / 2), kind="stable")]
        np.testing.assert_array_equal(got_rows[b, :len(order)].cpu().numpy(), perm[rows[b, order]])
    with pytest.raises(RuntimeError):
        hw.ops.rerank(t.cpu(), rows_t, "pair", anchor_rows=anchors_t)


def test_api_path_launches_no_torch_compute(c1_model, hw):
    """find_closest_neighbours_batch runs search, score convention and ordering in this library's kernels: the only
    torch kernels allowed on the path are copies / fills (host -> device uploads of row ids)."""
    from torch.profiler import ProfilerActivity, profile
    m, users = c1_model["gcn"], c1_model["users"]
    anchors = users[:64]
    m.find_closest_neighbours_batch("item", anchors, k=100)
    for model in (c1_model["gcn"], c1_model["base"]):
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            model.find_closest_neighbours_batch("item", anchors, k=100)
            torch.cuda.synchronize()
        names = [e.key for e in prof.key_averages() if getattr(e, "device_type", None) is not None
                 and str(e.device_type).endswith("CUDA")]
        bad = [x for x in names if x.startswith("void at::") or "at::native" in x]
        bad = [x for x in bad if not any(t in x.lower() for t in ("copy", "fill", "memcpy", "memset"))]
        assert not bad, bad
        assert any("rerank_kernel" in x for x in names), names


# ----------------------------------------------------------------------------- BASELINE.json full size (config C4)
def test_full_size_properties_10m(hw):
    """10M x 128, top-100 (the headline workload): properties that do not need a CPU pass over the table."""
    n, d, k, B = 10_000_000, 128, 100, 64
    gen = torch.Generator(device="cuda").manual_seed(0)
    table, shadow = hw.ops.blend_normalize(torch.randn((n, d), generator=gen, device="cuda"),
                                           torch.randn((n, d), generator=gen, device="cuda"), 0.5)
    q = hw.ops.unit_length(torch.randn((B, d), generator=gen, device="cuda"))
    # plant: query j's exact copy at row 1000 + 37 j, and a close second at the very last row of the table
    for j in range(4):
        table[1000 + 37 * j] = q[j]
    v = q[0] + 0.05 * hw.ops.unit_length(torch.randn((1, d), generator=gen, device="cuda"))[0]
    table[n - 1] = v / v.norm()
    shadow = hw.ops.make_shadow(table)
    v_, mean, _, _, mx = hw.ops.norm_stats(table)
    assert v_ == 0
    index = hw.ops.TopKIndex(table, shadow, max_norm=mx)
    idx, sc, s64 = index.topk(q, k, want_f64=True)
    assert bool((s64[:, :-1] >= s64[:, 1:]).all())                              # sortedness
    assert all(len(set(r)) == k for r in idx.cpu().tolist())                    # no duplicates
    for j in range(4):
        assert idx[j, 0].item() == 1000 + 37 * j and abs(s64[j, 0].item() - 1.0) < 1e-6
    assert idx[0, 1].item() == n - 1                                            # last row (tail tile) is reachable
    # re-scoring the returned rows reproduces the scores (fp64), and nothing outside beats the k-th score:
    rows = table[idx.reshape(-1)].double().reshape(B, k, d)
    np.testing.assert_allclose(torch.einsum("bkd,bd->bk", rows, q.double()).cpu().numpy(), s64.cpu().numpy(), atol=1e-12)
    for j in (0, 5, 63):
        # independent fp64 pass over the whole table on the GPU, in 1M-row chunks
        full = torch.cat([table[b:b + 1_000_000].double() @ q[j].double() for b in range(0, n, 1_000_000)])
        kth = s64[j, -1].item()
        # (the two fp64 sums use different summation orders: compare with a 1e-12 guard band)
        assert int((full > kth + 1e-12).sum().item()) <= k - 1 and int((full >= kth - 1e-12).sum().item()) >= k
        assert set(torch.topk(full, k).indices.cpu().tolist()) == set(idx[j].cpu().tolist())
    # idempotence + shard/merge equality at full size
    idx2, _, s642 = index.topk(q, k, want_f64=True)
    assert torch.equal(idx, idx2) and torch.equal(s64, s642)
    halves = []
    for g in range(2):
        b, e = hw.sharded.partition(n, 2, g)
        halves.append(hw.sharded.ShardedTopK(table[b:e], b, shadow=shadow[b:e], max_norm=mx).local_topk(q, k))
    midx, _, ms64 = hw.ops.merge_topk(torch.stack([h[1] for h in halves]).contiguous(),
                                      torch.stack([h[0] for h in halves]).contiguous(), want_f64=True)
    assert torch.equal(midx, idx) and torch.equal(ms64, s64)
    # bf16 mode on the same index: recall against the exact answer
    bidx, _ = index.topk(q, k, "bf16")
    rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(bidx.cpu().tolist(), idx.cpu().tolist())])
    assert rec >= 0.97

    # ---- the batches bench.py times: B = 4096 (16 query blocks, 8 filter launches, spill_cap 64) and B = 1, each
    # against an independent fp64 pass over the whole table, and against the B = 64 schedule's answer
    def full_pass(qv):
        return torch.cat([table[b:b + 1_000_000].double() @ qv.double() for b in range(0, n, 1_000_000)])

    q4k = torch.cat([q, hw.ops.unit_length(torch.randn((4096 - B, d), generator=gen, device="cuda"))])
    idx4, _, s4 = index.topk(q4k, k, want_f64=True)
    assert torch.equal(idx4[:B], idx) and torch.equal(s4[:B], s64)              # same rows as the B = 64 schedule
    assert bool((s4[:, :-1] >= s4[:, 1:]).all())
    srt = idx4.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                              # no duplicates in any of the 4096 lists
    for j in (64, 1999, 4095):
        full = full_pass(q4k[j])
        kth = s4[j, -1].item()
        assert int((full > kth + 1e-12).sum().item()) <= k - 1 and int((full >= kth - 1e-12).sum().item()) >= k
        assert set(torch.topk(full, k).indices.cpu().tolist()) == set(idx4[j].cpu().tolist())
        np.testing.assert_allclose(full[idx4[j]].cpu().numpy(), s4[j].cpu().numpy(), atol=1e-12)
    for j in (0, 777):
        idx1, _, s1 = index.topk(q4k[j:j + 1].contiguous(), k, want_f64=True)   # B = 1: the HBM-bound schedule
        assert torch.equal(idx1[0], idx4[j]) and torch.equal(s1[0], s4[j])
    # 2-shard and 8-shard merges of the B = 4096 batch equal the single-index answer bit for bit
    for G in (2, 8):
        parts = []
        for g in range(G):
            b, e = hw.sharded.partition(n, G, g)
            parts.append(hw.sharded.ShardedTopK(table[b:e], b, shadow=shadow[b:e], max_norm=mx).local_topk(q4k, k))
        midx, _, ms64 = hw.ops.merge_topk(torch.stack([h[1] for h in parts]).contiguous(),
                                          torch.stack([h[0] for h in parts]).contiguous(), want_f64=True)
        assert torch.equal(midx, idx4) and torch.equal(ms64, s4)
        del parts


# ----------------------------------------------------------------------------- GCN inference
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_gcn_infer_matches_the_reference_module(hw, golden_gcn, ci):
    """hwer_gcn_infer against outputs of the reference's GraphConvModule.forward (tests/golden/reference_gcn.npz,
    oracle/make_golden_gcn.py; 1 / 2 / 3 GCN layers, content widths 48 / 20 / 36): node vectors and the EMA state.
    fp32 tolerance 2e-6 (the kernels sum in a different order than torch's CPU GEMM)."""
    from conftest import gcn_case
    c = gcn_case(golden_gcn, ci)
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    prev = t(c["previous"])
    nbr = [(t(p, torch.int64), t(i, torch.int64)) for p, i in c["nbr"]]
    out = hw.ops.gcn_infer(t(c["node_emb"]), t(c["content"]), t(c["proj_w"]), t(c["proj_b"]), t(c["ln_g"]), t(c["ln_b"]),
                           nbr, t(c["fc0_w"]), t(c["fc0_b"]), t(c["fc1_w"]), t(c["fc1_b"]), previous=prev, ema=0.1)
    np.testing.assert_allclose(out.cpu().numpy(), c["h"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(prev.cpu().numpy(), c["previous_after"], rtol=0, atol=2e-6)


def test_gcn_vectors_api_matches_oracle_and_feeds_the_search(hw):
    """GcnNCF.get_gcn_vectors (neighbours drawn from the edge list by the host sampler, C = 50: not a multiple of 4,
    several node chunks' worth of rows is covered by the golden cases' logic) == oracle.gcn_infer on the same lists;
    and the vectors go through prepare_for_knn / __build_knn__ into a search like the reference's fit() does."""
    rs = np.random.RandomState(31)
    n_u, n_i, C, F, L = 300, 500, 50, 32, 2
    n = n_u + n_i
    users = [hw.Node("user", i) for i in range(n_u)]
    items = [hw.Node("item", i) for i in range(n_i)]
    edges = [hw.Edge(users[int(rs.randint(n_u))], items[int(rs.randint(n_i))], 1.0) for _ in range(4000)]
    params = {"node_emb": rs.standard_normal((n + 1, F)).astype(np.float32) / F,
              "proj_w": (rs.standard_normal((F, C)) * 0.2).astype(np.float32),
              "proj_b": (rs.standard_normal(F) * 0.01).astype(np.float32),
              "ln_g": (1 + 0.1 * rs.standard_normal(F)).astype(np.float32), "ln_b": (0.1 * rs.standard_normal(F)).astype(np.float32),
              "fc0_w": (rs.standard_normal((4 * F, F * (L + 1))) * 0.1).astype(np.float32),
              "fc0_b": (rs.standard_normal(4 * F) * 0.01).astype(np.float32),
              "fc1_w": (rs.standard_normal((F, 4 * F)) * 0.1).astype(np.float32),
              "fc1_b": (rs.standard_normal(F) * 0.01).astype(np.float32)}
    content = rs.standard_normal((n, C)).astype(np.float32)
    model = hw.GcnNCF(None, {"user", "item"}, n_dims=F)
    model.add_nodes(users + items)
    src = model.nodes_to_idx.rows_of([e.src for e in edges])
    dst = model.nodes_to_idx.rows_of([e.dst for e in edges])
    nbr = hw.ops.sample_neighbours(n, src, dst, fanout=2, seed=5, blocks=L)
    got = model.get_gcn_vectors(params, content, edges=edges, seed=5)
    want, _ = O.gcn_infer(params["node_emb"], content, params["proj_w"], params["proj_b"], params["ln_g"], params["ln_b"],
                          nbr, params["fc0_w"], params["fc0_b"], params["fc1_w"], params["fc1_b"], None)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-6)
    np.testing.assert_allclose(got.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)      # no EMA: unit rows
    # the same vectors, given explicitly, and into the serving table
    got2 = model.get_gcn_vectors(params, content, neighbours=nbr)
    assert torch.equal(got, got2)
    served = hw.GcnNCF(None, {"user", "item"}, n_dims=F)
    served.fit(users + items, edges, None, collaborative_vectors=got)
    res = served.find_closest_neighbours("item", users[3], k=10)
    assert len(res) == 10 and all(nd.node_type == "item" for nd, _ in res)

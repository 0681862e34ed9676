"""GPU (-m gpu): the callers around the retrieval path -- link-prediction metrics, the
validation harness and the single-list ranking metrics of hwer/utils.py -- against the oracle and against outputs of
the reference itself (tests/golden/reference_lp.npz, reference_eval.npz); plus the k = 1000 retrieval of config C5."""
import random

import numpy as np
import pytest
import torch

import hwer_oracle as O
from conftest import synthetic_case, synthetic_edges

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hw():
    import hwer_b200
    from hwer_b200 import _native
    assert _native.library_path().endswith("libhwer_b200.so")
    return hwer_b200


# ----------------------------------------------------------------------------- hwer_link_metrics
@pytest.mark.parametrize("P,levels,pos_rate", [(1, 2, 1.0), (2, 1, 0.5), (1000, 7, 0.3), (4097, 3, 0.1),
                                               (300001, 1 << 20, 0.09), (50000, 50, 0.0), (50000, 50, 1.0),
                                               (2_000_003, 1 << 24, 0.09)])
def test_link_metrics_matches_oracle(hw, P, levels, pos_rate):
    rs = np.random.RandomState(P % 9973)
    labels = (rs.random_sample(P) < pos_rate).astype(np.uint8)
    if 0.0 < pos_rate < 1.0:
        labels[0], labels[-1] = 1, 0
    scores = (rs.randint(0, levels, P).astype(np.float64) / levels).astype(np.float32)       # many exact ties
    if P > 3:
        scores[1], scores[2] = 0.0, -0.0                                                     # one threshold
    got = hw.ops.link_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda()).cpu().numpy()
    tp = float(np.sum((labels == 1) & (scores >= 0.5))); fp = float(np.sum((labels == 0) & (scores >= 0.5)))
    fn = float(np.sum((labels == 1) & (scores < 0.5))); tn = float(np.sum((labels == 0) & (scores < 0.5)))
    np.testing.assert_array_equal(got[4:], [tp, fp, fn, tn])                                 # counts: bit-exact
    if labels.sum() == 0:
        assert got[0] == 0.0 and got[2] == 0.0
        return
    ap, precision, recall, acc = O.link_prediction_metrics(labels, scores)
    np.testing.assert_allclose(got[:4], [ap, precision, recall, acc], rtol=1e-12, atol=1e-15)


def test_link_metrics_equals_sklearn(hw):
    from sklearn.metrics import accuracy_score, average_precision_score, precision_recall_fscore_support
    rs = np.random.RandomState(5)
    labels = (rs.random_sample(20000) < 0.2).astype(np.uint8)
    scores = np.clip(0.5 + 0.2 * rs.standard_normal(20000) + 0.1 * labels, 0, 1).astype(np.float32)
    scores[:2000] = np.round(scores[:2000], 2)
    got = hw.ops.link_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda()).cpu().numpy()
    p, r, _, _ = precision_recall_fscore_support(labels, scores >= 0.5, average="binary")
    want = [average_precision_score(labels, scores), p, r, accuracy_score(labels, scores >= 0.5)]
    np.testing.assert_allclose(got[:4], want, rtol=1e-12)


def test_link_metrics_rejects_cpu_and_empty(hw):
    with pytest.raises(RuntimeError):
        hw.ops.link_metrics(torch.zeros(4), torch.zeros(4, dtype=torch.uint8))
    with pytest.raises(ValueError):
        hw.ops.link_metrics(torch.zeros(0, device="cuda"), torch.zeros(0, dtype=torch.uint8, device="cuda"))


# ----------------------------------------------------------------------------- harness vs the reference's outputs
@pytest.fixture(scope="module")
def lp_case(hw, golden_lp):
    nu, ni, dd = [int(x) for x in golden_lp["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_lp["seeds"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(golden_lp["seeds"][1]))
    train = [hw.Edge(users[u], items[i], w) for u, i, w in tr]
    val = [hw.Edge(users[u], items[i], w) for u, i, w in vl]
    table = O.unit_length(collab, axis=1)
    m = hw.ContentRecommendation(None, {"user", "item"}, n_dims=dd)
    m.fit(users + items, train, None, vectors=table)
    return dict(model=m, nodes=users + items, train=train, val=val, table=table, dd=dd, g=golden_lp)


def test_link_prediction_accuracy_matches_reference(hw, lp_case):
    g = lp_case["g"]
    keys = [str(k) for k in g["lp_keys"]]
    # fp32 pair scores are summed in another order than numpy's: a pair within ~1e-7 of the 0.5 threshold, or of
    # a differently-labelled neighbour, may flip -> each such flip moves a metric by O(1/P), P ~ 7e4
    random.seed(int(g["seeds"][2]))
    got = hw.validation.link_prediction_accuracy(lp_case["model"], lp_case["nodes"], lp_case["train"], lp_case["val"])
    assert sorted(got.keys()) == sorted(keys)
    for k, v in zip(keys, g["lp_values"]):
        assert abs(got[k] - v) < 2e-4, (k, got[k], v)
    random.seed(int(g["seeds"][3]))
    got = hw.validation.link_prediction_accuracy(lp_case["model"], lp_case["nodes"], lp_case["train"][:500],
                                                 lp_case["val"][:50])
    for k, v in zip(keys, g["lp2_values"]):
        assert abs(got[k] - v) < 2e-3, (k, got[k], v)       # P = 550 validation pairs: one flip is 1.8e-3


def test_get_prediction_details_matches_reference(hw, lp_case):
    g = lp_case["g"]
    random.seed(int(g["seeds"][4]))
    preds, actuals, stats = hw.validation.get_prediction_details(lp_case["model"], lp_case["nodes"], lp_case["train"],
                                                                 lp_case["val"], hw.validation.model_get_topk, "item")
    np.testing.assert_allclose(preds, g["details_predictions"], atol=1e-6)
    np.testing.assert_array_equal(actuals, g["details_actuals"])
    assert sorted(stats.keys()) == [str(k) for k in g["details_keys"]]      # the reference's metric dict, key for key
    for k, v in zip([str(k) for k in g["ee_keys"]], g["details_ee_values"]):
        assert abs(stats[k] - v) < 1e-9, (k, stats[k], v)
    # ncf_eval draws the same NUMBER of random samples as the reference, so the link-prediction pair sets that follow
    # are the reference's own
    for k, v in zip([str(k) for k in g["lp_keys"]], g["details_lp_values"]):
        assert abs(stats[k] - v) < 2e-4, (k, stats[k], v)
    assert 0.0 <= stats["ncf_hr"] <= 1.0 and 0.0 <= stats["ncf_ndcg"] <= 1.0


def test_run_models_for_testing_end_to_end(hw, lp_case, capsys):
    dd, table = lp_case["dd"], lp_case["table"]
    edges = [(e, False) for e in lp_case["train"]] + [(e, True) for e in lp_case["val"]]
    hp = {"content": {"n_dims": dd, "vectors": table},
          "gcn_ncf": {"n_dims": dd, "collaborative_vectors": table, "content_vectors": None, "alpha": 0.0}}
    random.seed(1)
    ndcg, ncf_ndcg = hw.validation.run_models_for_testing(lp_case["nodes"], edges, {"user", "item"}, "item",
                                                          lambda: (None, None), ["content"], hp, display=False)
    ref = dict(zip([str(k) for k in lp_case["g"]["ee_keys"]], lp_case["g"]["details_ee_values"]))
    assert abs(ndcg - ref["ndcg_b@100"]) < 1e-9 and 0.0 <= ncf_ndcg <= 1.0
    ndcg2, _ = hw.validation.run_model_for_hpo(lp_case["nodes"], edges, {"user", "item"}, "item", lambda: (None, None),
                                               hp["gcn_ncf"], "gcn_ncf")
    assert abs(ndcg2 - ndcg) < 1e-9          # same table, same retrieval; only the score convention differs
    ndcg3, _ = hw.validation.run_models_for_testing(lp_case["nodes"], edges, {"user", "item"}, "item",
                                                    lambda: (None, None), ["content"], hp, display=True,
                                                    results_csv=None)
    assert abs(ndcg3 - ndcg) < 1e-9 and "retrieval_time" in capsys.readouterr().out


# ----------------------------------------------------------------------------- hwer/utils.py:71-121 one list at a time
def test_single_list_metrics_match_reference_spot_values(hw, golden_eval):
    ut = hw.utils
    y_true = {"a": 1, "b": 1, "c": 1}
    got = [ut.ndcg(y_true, ["x", "a", "b"]), ut.recall(y_true, ["x", "a", "b"]), ut.reciprocal_rank(["a"], ["x", "a"]),
           ut.binary_ndcg({"a": 5.0, "b": 2.0}, ["b", "q", "a"]),
           ut.ndcg({"s1": 5.0, "s2": 4.8, "s3": 3.0, "s4": 4.1, "s5": 2.9, "s6": 0.9}, ["s1", "s2", "s3", "s5", "s6"])]
    np.testing.assert_allclose(got, golden_eval["spot"], rtol=1e-6)       # relevances travel as fp32
    np.testing.assert_allclose(got[:4], golden_eval["spot"][:4], rtol=1e-12)
    rs = np.random.RandomState(0)
    for _ in range(20):
        n_true, n_pred = rs.randint(0, 12), rs.randint(0, 40)
        yt = {int(i): float(rs.randint(1, 6)) for i in rs.choice(60, n_true, replace=False)}
        yp = [int(i) for i in rs.choice(60, n_pred, replace=False)]
        assert abs(ut.ndcg(yt, yp) - O.ndcg(yt, yp)) < 1e-12
        assert abs(ut.binary_ndcg(yt, yp) - O.binary_ndcg(yt, yp)) < 1e-12
        assert abs(ut.binary_ndcg_v2(list(yt), yp) - O.binary_ndcg_v2(list(yt), yp)) < 1e-12
        assert abs(ut.recall(yt, yp) - O.recall(yt, yp)) < 1e-12
        assert abs(ut.reciprocal_rank(list(yt), yp) - O.reciprocal_rank(list(yt), yp)) < 1e-12
    assert abs(ut.average_precision(["a", "b"], ["x", "a", "a", "b"]) - (1 / 2 + 2 / 4) / 2) < 1e-12


# ----------------------------------------------------------------------------- k = 1000 (config C5's k, scaled rows)
@pytest.mark.parametrize("n,d,B,k", [(200000, 128, 64, 1000), (60000, 128, 520, 1000), (5000, 64, 3, 1000)])
def test_topk_1000_matches_oracle(hw, n, d, B, k):
    rs = np.random.RandomState(n % 1000 + B)
    t_np = O.unit_length(rs.standard_normal((n, d)).astype(np.float32), axis=1)
    q_np = O.unit_length(rs.standard_normal((B, d)).astype(np.float32), axis=1)
    t, q = torch.from_numpy(t_np).cuda(), torch.from_numpy(q_np).cuda()
    idx, sc, s64 = hw.ops.TopKIndex(t).topk(q, k, "exact", want_f64=True)
    ref_idx, ref_sc = O.exact_topk(t_np, q_np, k)
    idx, sc, s64 = idx.cpu().numpy(), sc.cpu().numpy(), s64.cpu().numpy()
    assert O.compare_topk(idx, s64, ref_idx, ref_sc, tie_eps=1e-6) == 0
    assert (idx == ref_idx).mean() > 0.999
    np.testing.assert_allclose(s64, ref_sc, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sc, ref_sc, rtol=0, atol=1e-5)
    # shard / merge equality at k = 1000 (8 shards: the C5 layout)
    parts = []
    for g in range(8):
        b, e = hw.sharded.partition(n, 8, g)
        if e - b < k:
            return
        parts.append(hw.sharded.ShardedTopK(t[b:e].contiguous(), b).local_topk(q, k))
    midx, _, ms64 = hw.ops.merge_topk(torch.stack([p[1] for p in parts]).contiguous(),
                                      torch.stack([p[0] for p in parts]).contiguous(), want_f64=True)
    np.testing.assert_array_equal(midx.cpu().numpy(), idx)
    np.testing.assert_array_equal(ms64.cpu().numpy(), s64)


def test_query_sharded_single_process_equals_plain_index(hw):
    rs = np.random.RandomState(4)
    t = torch.from_numpy(O.unit_length(rs.standard_normal((27278, 256)).astype(np.float32), axis=1)).cuda()   # C3 items
    q = torch.from_numpy(O.unit_length(rs.standard_normal((700, 256)).astype(np.float32), axis=1)).cuda()
    a = hw.sharded.QueryShardedTopK(t).topk(q, 100)
    b = hw.ops.TopKIndex(t).topk(q, 100)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert hw.sharded.QueryShardedTopK(t).local_slice(700) == (0, 700)


def test_c5_shard_properties_62m_top1000(hw):
    """Config C5 as one of its eight GPUs sees it: a 62.5 M x 128 shard (32 GB fp32 + 16 GB bf16 shadow), exact
    top-1000, batch 4096, rows offset to the shard's global ids.  Size-independent properties only."""
    n, d, k, B, shard = 62_500_000, 128, 1000, 4096, 7
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * (1 << 30):
        pytest.skip("needs ~60 GB of free HBM")
    gen = torch.Generator(device="cuda").manual_seed(5)
    table = torch.empty((n, d), dtype=torch.float32, device="cuda")
    step = 4_000_000
    for b in range(0, n, step):
        e = min(n, b + step)
        table[b:e] = hw.ops.unit_length(torch.randn((e - b, d), generator=gen, device="cuda"))
    q = hw.ops.unit_length(torch.randn((B, d), generator=gen, device="cuda"))
    for j in range(3):                                   # plant exact copies: first tile, middle, the very last row
        table[(5, n // 2 + 11, n - 1)[j]] = q[j]
    shadow = hw.ops.make_shadow(table)
    v_, _, _, _, mx = hw.ops.norm_stats(table)
    assert v_ == 0
    index = hw.ops.TopKIndex(table, shadow, max_norm=mx)
    off = shard * n
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    index.topk(q[:256].contiguous(), k, idx_offset=off)  # sizes the workspace
    t0.record()
    idx, sc, s64 = index.topk(q, k, idx_offset=off, want_f64=True)
    t1.record()
    torch.cuda.synchronize()
    print("C5 shard: %d x %d, top-%d, batch %d: %.1f ms" % (n, d, k, B, t0.elapsed_time(t1)))
    assert bool((s64[:, :-1] >= s64[:, 1:]).all())                                   # sortedness
    assert int(idx.min().item()) >= off and int(idx.max().item()) < off + n          # global ids of this shard
    srt = idx.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                                   # no duplicates in any row
    for j in range(3):
        assert idx[j, 0].item() == off + (5, n // 2 + 11, n - 1)[j] and abs(s64[j, 0].item() - 1.0) < 1e-6
    # returned scores are the fp64 dot products of the returned rows
    sub = slice(0, 64)
    rows = table[(idx[sub] - off).reshape(-1)].double().reshape(64, k, d)
    np.testing.assert_allclose(torch.einsum("bkd,bd->bk", rows, q[sub].double()).cpu().numpy(),
                               s64[sub].cpu().numpy(), atol=1e-12)
    # nothing outside the result beats its k-th score: independent fp64 pass over the whole shard
    for j in (0, 1000, 4095):
        full = torch.cat([table[b:b + step].double() @ q[j].double() for b in range(0, n, step)])
        kth = s64[j, -1].item()
        assert int((full > kth + 1e-12).sum().item()) <= k - 1 and int((full >= kth - 1e-12).sum().item()) >= k
        assert set(torch.topk(full, k).indices.cpu().tolist()) == set((idx[j] - off).cpu().tolist())
        del full
    # batch independence: a query answered alone gives the same row
    i1, _, s1 = index.topk(q[1000:1001].contiguous(), k, idx_offset=off, want_f64=True)
    assert torch.equal(i1[0], idx[1000]) and torch.equal(s1[0], s64[1000])


# ----------------------------------------------------------------------------- BASELINE.json configs[1] / configs[2]
def test_c2_extraction_efficiency_matches_reference_run(hw, golden_c2):
    """ML-1M shape, 6,040 x 3,706, d = 128: the whole validation.extraction_efficiency (top-200 for every user,
    train items filtered, Recall@K / NDCG / diversity) against the metrics the reference produced on the same
    seeded inputs (58.7 s there, hwer/validation.py:100-187)."""
    import time
    nu, ni, dd = [int(x) for x in golden_c2["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_c2["seeds"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(golden_c2["seeds"][1]))
    train = [hw.Edge(users[u], items[i], w) for u, i, w in tr]
    val = [hw.Edge(users[u], items[i], w) for u, i, w in vl]
    m = hw.ContentRecommendation(None, {"user", "item"}, n_dims=dd)
    m.fit(users + items, train, None, vectors=O.unit_length(collab, axis=1))
    random.seed(0)
    t0 = time.time()
    res = hw.validation.extraction_efficiency(m, train, val, hw.validation.model_get_topk, "item")
    print("C2 extraction_efficiency: %.2f s (retrieval %.4f s; the reference: 58.7 s / %.1f s)"
          % (time.time() - t0, res["metrics"]["retrieval_time"], float(golden_c2["retrieval_time"][0])))
    ref = dict(zip([str(k) for k in golden_c2["metric_keys"]], golden_c2["metric_values"]))
    for key, v in ref.items():
        assert abs(res["metrics"][key] - v) < 1e-9, (key, res["metrics"][key], v)
    # a few users' filtered top-100 lists, id for id
    rows = res["rows"].cpu().numpy()
    pos = {u: j for j, u in enumerate(res["users"])}
    train_items = {}
    for u, i, w in tr:
        train_items.setdefault(u, set()).add(i)
    for u, want in zip(golden_c2["sample_users"], golden_c2["sample_preds"]):
        got = [int(r) - nu for r in rows[pos[users[int(u)]]] if int(r) - nu not in train_items.get(int(u), ())][:100]
        assert got == [int(x) for x in want if x >= 0], int(u)


def test_c3_find_closest_neighbours_matches_reference_run(hw, golden_c3):
    """ML-20M item side, 27,278 x 256: top-100 for 96 user and 16 item anchors, ids and scores of the reference."""
    nu, ni, dd, k = [int(x) for x in golden_c3["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_c3["seed"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    edges = [hw.Edge(users[i % nu], items[i], 1.0) for i in range(10)]
    m = hw.ContentRecommendation(None, {"user", "item"}, n_dims=dd)
    m.fit(users + items, edges, None, vectors=O.unit_length(collab, axis=1))
    for anchors, nodes, key in ((golden_c3["user_anchors"], users, "user"), (golden_c3["item_anchors"], items, "item")):
        rows, scores = m.find_closest_neighbours_batch("item", [nodes[int(a)] for a in anchors], k=k)
        np.testing.assert_array_equal(rows.cpu().numpy() - nu, golden_c3[key + "_idx"])
        np.testing.assert_allclose(scores.cpu().numpy(), golden_c3[key + "_score"], rtol=0, atol=1e-6)
    one = m.find_similar_items(items[int(golden_c3["item_anchors"][0])], k=k)
    assert [int(n.node_external_id) for n, s in one] == [int(x) for x in golden_c3["item_idx"][0]]

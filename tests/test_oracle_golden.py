"""CPU: the oracle (oracle/hwer_oracle.py) against outputs of the REFERENCE ITSELF (tests/golden/*.npz, made by
oracle/make_golden.py from /root/reference).  This is what pins the oracle."""
import os

import numpy as np
import pytest

import hwer_oracle as O
from conftest import GOLDEN, posneg_lists, synthetic_case, synthetic_edges


@pytest.fixture(scope="module")
def c1(golden_c1):
    n_users, n_items, d, k = [int(x) for x in golden_c1["shape"]]
    content, collab = synthetic_case(n_users, n_items, d, seed=int(golden_c1["seed"][0]))
    table = O.blend_normalize(content, collab, 0.0)        # alpha = 0 == reference prepare_for_knn
    users = [O.Node("user", i) for i in range(n_users)]
    items = [O.Node("item", i) for i in range(n_items)]
    r = O.OracleRecommender({"user", "item"}, n_dims=d)
    r.add_nodes(users + items)
    r.build_knn(table)
    return dict(g=golden_c1, table=table, users=users, items=items, rec=r, k=k, content=content, collab=collab)


def test_table_matches_reference_prepare_for_knn(c1):
    g, table = c1["g"], c1["table"]
    assert table.dtype == np.float32
    np.testing.assert_array_equal(table[[0, 1, 942, 943, 2624]], g["table_rows"])
    assert abs(float(np.abs(table.astype(np.float64)).sum()) - float(g["table_checksum"][0])) < 1e-6


def test_unit_length_violations(c1):
    g, table = c1["g"], c1["table"]
    np.testing.assert_allclose([float(x) for x in O.unit_length_violations(table, axis=1)], g["viol"], rtol=1e-12)
    p = table.copy()
    p[3] *= 1.01
    p[10] *= 0.9
    p[11] *= 1.0 + 5e-5
    got = [float(x) for x in O.unit_length_violations(p, axis=1)]
    np.testing.assert_allclose(got, g["viol_perturbed"], rtol=1e-12)
    assert got[0] == 2 and got[2] == 1 and got[3] == 1


def _ids(res):
    return [int(n.node_external_id) for n, s in res], [float(s) for n, s in res]


def test_find_items_for_user_matches_reference(c1):
    g, r = c1["g"], c1["rec"]
    for j, u in enumerate(g["user_anchors"]):
        idx, sc = _ids(r.find_closest_neighbours("item", c1["users"][int(u)], k=c1["k"]))
        assert idx == list(g["items_for_user_idx"][j])
        np.testing.assert_allclose(sc, g["items_for_user_score"][j], rtol=0, atol=1e-7)


def test_find_similar_items_matches_reference(c1):
    g, r = c1["g"], c1["rec"]
    for j, i in enumerate(g["item_anchors"]):
        idx, sc = _ids(r.find_closest_neighbours("item", c1["items"][int(i)], k=c1["k"]))
        assert idx == list(g["similar_items_idx"][j])
        assert idx[0] == int(i)          # the anchor item itself comes back first
        np.testing.assert_allclose(sc, g["similar_items_score"][j], rtol=0, atol=1e-7)


def test_default_k_and_posneg_match_reference(c1):
    g, r = c1["g"], c1["rec"]
    users, items = c1["users"], c1["items"]
    n_items = len(items)
    for j, u in enumerate(g["user_anchors"][:8]):
        idx, sc = _ids(r.find_closest_neighbours("user", users[int(u)]))      # default k = 200
        assert len(idx) == 200 and idx == list(g["users_k200_idx"][j])
    for j, u in enumerate(g["user_anchors"][:16]):
        u = int(u)
        pos = [items[(u * 3 + t) % n_items] for t in range(3)]
        neg = [items[(u * 5 + t + 1) % n_items] for t in range(2)]
        idx, sc = _ids(r.find_closest_neighbours("item", users[u], positive=pos, negative=neg, k=c1["k"]))
        assert idx == list(g["posneg_idx"][j])
        np.testing.assert_allclose(sc, g["posneg_score"][j], atol=1e-7)


def test_gcn_score_convention_matches_reference(c1):
    g = c1["g"]
    r = O.OracleRecommender({"user", "item"}, n_dims=64, gcn_scores=True)
    r.add_nodes(c1["users"] + c1["items"])
    r.build_knn(c1["table"])
    for j, u in enumerate(g["user_anchors"]):
        idx, sc = _ids(r.find_closest_neighbours("item", c1["users"][int(u)], k=c1["k"]))
        assert idx == list(g["gcn_items_for_user_idx"][j])
        np.testing.assert_allclose(sc, g["gcn_items_for_user_score"][j], atol=1e-12)


def test_predict_matches_reference_incl_unknown_nodes(c1):
    g, r = c1["g"], c1["rec"]
    nodes = c1["users"] + c1["items"]
    pairs = []
    for a, b in zip(g["pair_src"], g["pair_dst"]):
        pairs.append((nodes[a] if a >= 0 else O.Node("user", "ghost%d" % len(pairs)),
                      nodes[b] if b >= 0 else O.Node("item", "ghost%d" % len(pairs))))
    np.testing.assert_allclose(r.predict(pairs), g["pair_pred"], atol=1e-7)
    assert abs(g["pair_pred"][-1] - 0.5) < 1e-6          # unknown x unknown ~ 0.5


def test_exact_topk_equals_reference_order(c1):
    """Brute-force fp64 (score desc, row asc) == the reference's KD-tree retrieval on this table."""
    g, table = c1["g"], c1["table"]
    n_users = len(c1["users"])
    q = O.unit_length(table[g["user_anchors"]], axis=1)
    idx, sc = O.exact_topk(table[n_users:], q, c1["k"])
    np.testing.assert_array_equal(idx, g["items_for_user_idx"])
    np.testing.assert_allclose((sc + 1) / 2, g["items_for_user_score"], atol=1e-6)


def test_unknown_anchor_and_k_too_large(c1):
    r = c1["rec"]
    with pytest.raises(O.NodeNotFoundException):
        r.find_closest_neighbours("item", O.Node("user", "nobody"))
    with pytest.raises(ValueError):
        r.find_closest_neighbours("item", c1["users"][0], k=len(c1["items"]) + 1)


def test_metric_spot_values(golden_eval):
    y = {"a": 1, "b": 1, "c": 1}
    got = [O.ndcg(y, ["x", "a", "b"]), O.recall(y, ["x", "a", "b"]), O.reciprocal_rank(["a"], ["x", "a"]),
           O.binary_ndcg({"a": 5.0, "b": 2.0}, ["b", "q", "a"]),
           O.ndcg({"s1": 5.0, "s2": 4.8, "s3": 3.0, "s4": 4.1, "s5": 2.9, "s6": 0.9}, ["s1", "s2", "s3", "s5", "s6"])]
    np.testing.assert_allclose(got, golden_eval["spot"], rtol=1e-14)
    assert abs(got[0] - 0.5307212714866815) < 1e-12      # SURVEY.md section 8c spot value


def test_extraction_metrics_match_reference(golden_eval):
    nu, ni, dd = [int(x) for x in golden_eval["shape"]]
    _, collab = synthetic_case(nu, ni, dd, seed=int(golden_eval["seeds"][0]))
    table = O.unit_length(collab, axis=1)
    users = [O.Node("user", i) for i in range(nu)]
    items = [O.Node("item", i) for i in range(ni)]
    m = O.OracleRecommender({"user", "item"}, n_dims=dd)
    m.add_nodes(users + items)
    m.build_knn(table)
    tr, vl = synthetic_edges(nu, ni, seed=int(golden_eval["seeds"][1]))
    train = [(users[u], items[i], w) for u, i, w in tr]
    val = [(users[u], items[i], w) for u, i, w in vl]
    all_users = list(set([u for u, i, r in train] + [u for u, i, r in val]))
    preds = O.model_get_topk_knn(m, all_users, "item")
    out = O.extraction_metrics(preds, train, val, "item")
    ref = dict(zip([str(k) for k in golden_eval["metric_keys"]], golden_eval["metric_values"]))
    for key in ("recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"):
        assert abs(out[key] - ref[key]) < 1e-12, key


def test_compare_topk_tie_rule():
    ref_idx = np.array([[5, 3, 9, 1]])
    ref_sc = np.array([[0.9, 0.5, 0.5, 0.1]])
    assert O.compare_topk(np.array([[5, 9, 3, 1]]), ref_sc, ref_idx, ref_sc) == 0     # swap inside a tie group
    assert O.compare_topk(np.array([[3, 5, 9, 1]]), ref_sc, ref_idx, ref_sc) == 1     # real reordering
    assert O.compare_topk(np.array([[5, 3, 9, 7]]), ref_sc, ref_idx, ref_sc) == 0     # tie straddling the cut


# ----------------------------------------------------------------------------- NCF re-rank
@pytest.fixture(scope="module")
def golden_ncf():
    return np.load(os.path.join(GOLDEN, "reference_ncf.npz"))


@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_ncf_forward_matches_reference(golden_ncf, depth):
    """The oracle's restatement of hwer/ncf.py against the reference module's own outputs (its own initialisation)."""
    g = golden_ncf
    got = O.ncf_forward(g["h"], g["src"], g["dst"], g["params_d%d" % depth], depth)
    np.testing.assert_allclose(got, g["forward_d%d" % depth], rtol=0, atol=2e-6)


def test_ncf_predict_indexing_matches_reference(golden_ncf):
    """GcnNCF.predict's row convention (gcn_ncf.py:341-342): node row + 1, unknown node -> padding row 0."""
    g = golden_ncf
    got = O.ncf_forward(g["h"], g["predict_src"] + 1, g["predict_dst"] + 1, g["params_d3"], 3)
    np.testing.assert_allclose(got, g["predict"], rtol=0, atol=2e-6)


# ----------------------------------------------------------------------------- link prediction (validation.py:41-65)
def _eval_case(golden):
    nu, ni, dd = [int(x) for x in golden["shape"]]
    _, collab = synthetic_case(nu, ni, dd, seed=int(golden["seeds"][0]))
    users = [O.Node("user", i) for i in range(nu)]
    items = [O.Node("item", i) for i in range(ni)]
    m = O.OracleRecommender({"user", "item"}, n_dims=dd)
    m.add_nodes(users + items)
    m.build_knn(O.unit_length(collab, axis=1))
    tr, vl = synthetic_edges(nu, ni, seed=int(golden["seeds"][1]))
    return m, users + items, [(users[u], items[i], w) for u, i, w in tr], [(users[u], items[i], w) for u, i, w in vl]


def test_link_prediction_accuracy_matches_reference(golden_lp):
    import random
    m, nodes, train, val = _eval_case(golden_lp)
    keys = [str(k) for k in golden_lp["lp_keys"]]
    random.seed(int(golden_lp["seeds"][2]))
    got = O.link_prediction_accuracy(m, nodes, train, val, random)
    for k, v in zip(keys, golden_lp["lp_values"]):
        assert abs(got[k] - v) < 1e-12, (k, got[k], v)
    random.seed(int(golden_lp["seeds"][3]))
    got = O.link_prediction_accuracy(m, nodes, train[:500], val[:50], random)
    for k, v in zip(keys, golden_lp["lp2_values"]):
        assert abs(got[k] - v) < 1e-12, (k, got[k], v)


def test_average_precision_restatement_equals_sklearn_with_ties():
    from sklearn.metrics import average_precision_score
    rs = np.random.RandomState(3)
    for n, levels in ((1000, 7), (5000, 100000), (10, 2), (257, 1)):
        y = rs.randint(0, 2, n)
        y[0] = 1
        s = rs.randint(0, levels, n).astype(np.float32) / levels
        assert abs(O.average_precision_score(y, s) - average_precision_score(y, s)) < 1e-12


# ----------------------------------------------------------------------------- configs C2 / C3 (BASELINE.json)
def test_c2_extraction_metrics_match_reference(golden_c2):
    """ML-1M shape (6,040 x 3,706, d = 128): the oracle's restatement of validation.extraction_efficiency against the
    reference's own run (58.7 s there; candidates come from the brute-force exact_topk here, which
    test_exact_topk_equals_reference_order ties to the KD-tree's answers)."""
    nu, ni, dd = [int(x) for x in golden_c2["shape"]]
    _, collab = synthetic_case(nu, ni, dd, seed=int(golden_c2["seeds"][0]))
    table = O.unit_length(collab, axis=1)
    users = [O.Node("user", i) for i in range(nu)]
    items = [O.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, seed=int(golden_c2["seeds"][1]))
    train = [(users[u], items[i], w) for u, i, w in tr]
    val = [(users[u], items[i], w) for u, i, w in vl]
    idx, sc = O.exact_topk(table[nu:], table[:nu], 200)
    preds = {users[u]: [(items[int(i)], (float(s) + 1) / 2) for i, s in zip(idx[u], sc[u])] for u in range(nu)}
    out = O.extraction_metrics(preds, train, val, "item", cutoffs=(10, 100))
    ref = dict(zip([str(k) for k in golden_c2["metric_keys"]], golden_c2["metric_values"]))
    for key in ("recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"):
        assert abs(out[key] - ref[key]) < 1e-12, (key, out[key], ref[key])


def test_c3_find_closest_neighbours_matches_reference(golden_c3):
    """ML-20M item side (27,278 x 256): top-100 for user and item anchors, ids and scores of the reference's run."""
    nu, ni, dd, k = [int(x) for x in golden_c3["shape"]]
    _, collab = synthetic_case(nu, ni, dd, seed=int(golden_c3["seed"][0]))
    table = O.unit_length(collab, axis=1)
    users = [O.Node("user", i) for i in range(nu)]
    items = [O.Node("item", i) for i in range(ni)]
    m = O.OracleRecommender({"user", "item"}, n_dims=dd)
    m.add_nodes(users + items)
    m.build_knn(table)
    for anchors, nodes, key in ((golden_c3["user_anchors"][:24], users, "user"), (golden_c3["item_anchors"][:8], items, "item")):
        for j, a in enumerate(anchors):
            got = m.find_closest_neighbours("item", nodes[int(a)], k=k)
            ids = [int(n.node_external_id) for n, s in got]
            assert ids == [int(x) for x in golden_c3[key + "_idx"][j]]
            np.testing.assert_allclose([s for n, s in got], golden_c3[key + "_score"][j], rtol=0, atol=1e-6)
    # the brute-force order is the KD-tree's order (no ties at this shape)
    ua = golden_c3["user_anchors"]
    idx, sc = O.exact_topk(table[nu:], table[ua], k)
    np.testing.assert_array_equal(idx, golden_c3["user_idx"])


# ----------------------------------------------------------------------------- round 2 fixtures (reference_r2.npz)
def test_gcn_posneg_scores_are_distances_to_the_composed_embedding(c1, golden_r2):
    """GcnNCF.find_closest_neighbours with positive / negative lists (hwer/gcn_ncf.py:363-383)."""
    g, items, users = c1["g"], c1["items"], c1["users"]
    r = O.OracleRecommender({"user", "item"}, n_dims=c1["table"].shape[1], gcn_scores=True)
    r.add_nodes(users + items)
    r.build_knn(c1["table"])
    for j, u in enumerate(g["user_anchors"][:16]):
        p, n = posneg_lists(int(u), len(items))
        idx, sc = _ids(r.find_closest_neighbours("item", users[int(u)], [items[t] for t in p], [items[t] for t in n],
                                                 k=c1["k"]))
        assert idx == list(golden_r2["gcn_posneg_idx"][j])
        np.testing.assert_allclose(sc, golden_r2["gcn_posneg_score"][j], rtol=0, atol=1e-9)
    j = 0
    for u in g["user_anchors"][16:20]:
        p, n = posneg_lists(int(u), len(items))
        for pos, neg in (([items[t] for t in p], None), (None, [items[t] for t in n])):
            idx, sc = _ids(r.find_closest_neighbours("item", users[int(u)], pos, neg))
            assert idx == list(golden_r2["gcn_one_sided_idx"][j])
            np.testing.assert_allclose(sc, golden_r2["gcn_one_sided_score"][j], rtol=0, atol=1e-9)
            j += 1


def test_knn_query_and_embeddings_match_reference(c1, golden_r2):
    r, items, users = c1["rec"], c1["items"], c1["users"]
    u0 = int(golden_r2["knn_query_user"][0])
    p, n = posneg_lists(u0, len(items))
    emb = r.query_embedding(users[u0], [items[t] for t in p], [items[t] for t in n])
    res = r.knn.query(emb, "item", k=25)
    assert [int(x.node_external_id) for x, _ in res] == list(golden_r2["knn_query_idx"])
    np.testing.assert_allclose([d for _, d in res], golden_r2["knn_query_dist"], rtol=0, atol=1e-12)
    probe = [users[3], items[7], O.Node("user", "ghost"), items[1681], O.Node("item", "ghost2")]
    np.testing.assert_array_equal(r.get_embeddings(probe), golden_r2["emb_rows"])
    np.testing.assert_array_equal(r.get_average_embeddings([items[1], items[2], items[3]]), golden_r2["avg_a"])
    np.testing.assert_array_equal(r.get_average_embeddings([users[5], O.Node("item", "ghost3")]), golden_r2["avg_b"])


def eval_case(mod, golden_eval):
    """The 300 x 500 evaluation case of oracle/make_golden.py with `mod`'s Node type: (table, users, items, train, val)."""
    nu, ni, dd = [int(x) for x in golden_eval["shape"]]
    _, collab = synthetic_case(nu, ni, dd, int(golden_eval["seeds"][0]))
    users = [mod.Node("user", i) for i in range(nu)]
    items = [mod.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(golden_eval["seeds"][1]))
    return O.unit_length(collab, axis=1), users, items, tr, vl


def test_ncf_eval_matches_reference(golden_eval, golden_r2):
    import random
    table, users, items, tr, vl = eval_case(O, golden_eval)
    m = O.OracleRecommender({"user", "item"}, n_dims=table.shape[1])
    m.add_nodes(users + items)
    m.build_knn(table)
    train = [(users[u], items[i], w) for u, i, w in tr]
    val = [(users[u], items[i], w) for u, i, w in vl]
    all_items = [x for x in set([i for u, i, w in val] + [i for u, i, w in train]) if x.node_type == "item"]
    random.seed(int(golden_r2["ncf_eval_seed"][0]))
    hr, nd, _ = O.ncf_eval(m, train, val, all_items, random)
    np.testing.assert_allclose([hr, nd], golden_r2["ncf_eval"], rtol=0, atol=1e-12)


def test_prepare_for_knn_pca_branch_matches_reference(golden_r2):
    rs = np.random.RandomState(700)
    wide = (rs.standard_normal((400, 96)) * np.linspace(3.0, 0.2, 96)[None, :]).astype(np.float32)
    got = O.prepare_for_knn(wide, 32)
    assert got.shape == (400, 32)
    # sklearn decomposes the fp32 table in fp32: axes with close eigenvalues rotate by ~1e-4 against the float64
    # decomposition, while the Gram matrix -- all that retrieval sees -- agrees to 1e-5
    np.testing.assert_allclose(got, golden_r2["pca_table"], rtol=0, atol=3e-4)
    np.testing.assert_allclose(got @ got.T, golden_r2["pca_table"] @ golden_r2["pca_table"].T, rtol=0, atol=2e-5)
    with pytest.raises(ValueError):
        O.prepare_for_knn(wide[:, :16], 32)


# ----------------------------------------------------------------------------- GCN inference
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_gcn_infer_restatement_matches_the_reference_module(golden_gcn, ci):
    """oracle.gcn_infer against outputs of the reference's own GraphConvModule.forward (hwer/gcn.py:162-193, 1 / 2 / 3
    GCN layers), run in the build container on a stand-in NodeFlow with explicit neighbour lists
    (oracle/make_golden_gcn.py): node vectors and the updated EMA state."""
    from conftest import gcn_case
    c = gcn_case(golden_gcn, ci)
    out, prev = O.gcn_infer(c["node_emb"], c["content"], c["proj_w"], c["proj_b"], c["ln_g"], c["ln_b"], c["nbr"],
                            c["fc0_w"], c["fc0_b"], c["fc1_w"], c["fc1_b"], c["previous"], 0.1)
    np.testing.assert_allclose(out, c["h"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(prev, c["previous_after"], rtol=0, atol=1e-6)
    n = c["shape"][0]
    # rows are unit length before the EMA mixes 10 % of the previous state in: norms stay within that band
    nrm = np.linalg.norm(out, axis=1)
    assert nrm.max() < 1.0 + 0.1 * np.linalg.norm(c["previous"][:n], axis=1).max() + 1e-5

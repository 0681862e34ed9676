"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference/hwer, loaded through oracle/ref_shim.py) on seeded synthetic inputs.

Run in the build container (the only place /root/reference exists):   python oracle/make_golden.py
The fixtures are committed; tests regenerate the inputs from the same seeds (see `synthetic_case`).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def synthetic_case(n_users, n_items, d, seed):
    """Seeded inputs shared by this generator and the tests (legacy RandomState: stable across numpy versions)."""
    rs = np.random.RandomState(seed)
    content = rs.standard_normal((n_users + n_items, d)).astype(np.float32)
    collab = rs.standard_normal((n_users + n_items, d)).astype(np.float32)
    return content, collab


def synthetic_edges(n_users, n_items, seed, val_per_user=3, min_train=5, max_train=40):
    rs = np.random.RandomState(seed)
    train, val = [], []
    for u in range(n_users):
        deg = rs.randint(min_train, max_train + 1)
        items = rs.choice(n_items, size=deg + val_per_user, replace=False)
        ratings = rs.randint(1, 6, size=deg + val_per_user)
        for it, r in zip(items[:deg], ratings[:deg]):
            train.append((u, int(it), float(r)))
        if u % 7 != 0:   # some users have no validation edges
            for it, r in zip(items[deg:], ratings[deg:]):
                val.append((u, int(it), float(r)))
            if u % 5 == 0:   # a validation item that is also a train item (must be filtered)
                val.append((u, int(items[0]), 5.0))
    return train, val


def main():
    ref = ref_shim.load_reference()
    rb, ut, va, gn = ref.recommendation_base, ref.utils, ref.validation, ref.gcn_ncf
    os.makedirs(OUT, exist_ok=True)
    Node, Edge = rb.Node, rb.Edge

    class Dummy(rb.RecommendationBase):
        def fit(self, *a, **k):
            pass

    # ------------------------------------------------------------------ case 1: ML-100K shape (config C1)
    n_users, n_items, d, k = 943, 1682, 64, 10
    content, collab = synthetic_case(n_users, n_items, d, seed=100)
    users = [Node("user", i) for i in range(n_users)]
    items = [Node("item", i) for i in range(n_items)]
    # the reference's own table producer for the blend slot (alpha = 0 behaviour)
    g = gn.GcnNCF.__new__(gn.GcnNCF)
    g.n_dims = d
    table = g.prepare_for_knn(content, collab)                       # gcn_ncf.py:447-456
    assert table.dtype == np.float32
    viol = ut.unit_length_violations(table, axis=1)
    perturbed = table.copy()
    perturbed[3] *= 1.01
    perturbed[10] *= 0.9
    perturbed[11] *= 1.0 + 5e-5
    viol_p = ut.unit_length_violations(perturbed, axis=1)

    r = Dummy({"user", "item"}, n_dims=d)
    r.add_nodes(users + items)
    r.__build_knn__(table)
    r.fit_done = True
    rs = np.random.RandomState(7)
    user_anchors = rs.choice(n_users, 64, replace=False)
    item_anchors = rs.choice(n_items, 16, replace=False)

    def run(model, node_type, anchor, kk, pos=None, neg=None):
        res = model.find_closest_neighbours(node_type, anchor, positive=pos, negative=neg, k=kk)
        return ([int(n.node_external_id) for n, s in res], [float(s) for n, s in res])

    fi_idx, fi_sc = zip(*[run(r, "item", users[u], k) for u in user_anchors])          # find_items_for_user
    fs_idx, fs_sc = zip(*[run(r, "item", items[i], k) for i in item_anchors])          # find_similar_items
    fu_idx, fu_sc = zip(*[run(r, "user", users[u], 200) for u in user_anchors[:8]])    # default k, user side
    pn_idx, pn_sc = zip(*[run(r, "item", users[u], k, pos=[items[(u * 3 + j) % n_items] for j in range(3)],
                              neg=[items[(u * 5 + j + 1) % n_items] for j in range(2)]) for u in user_anchors[:16]])

    # GcnNCF serving override, cosine branch (gcn_ncf.py:363-383)
    g.node_types = {"user", "item"}
    g.nodes_to_idx = r.nodes_to_idx
    g.knn = r.knn
    g.vectors = r.vectors
    g.fit_done = True
    g.ncf_enabled = False
    gi_idx, gi_sc = zip(*[run(g, "item", users[u], k) for u in user_anchors])

    # predict incl. nodes never trained on (recommendation_base.py:135-151)
    pairs = [(users[int(a)], items[int(b)]) for a, b in zip(rs.randint(0, n_users, 200), rs.randint(0, n_items, 200))]
    pairs += [(Node("user", "ghost"), items[5]), (users[4], Node("item", "ghost")), (Node("user", "g1"), Node("item", "g2"))]
    pred = np.asarray(r.predict(pairs))
    pair_src = [r.nodes_to_idx.get(a, -1) for a, b in pairs]
    pair_dst = [r.nodes_to_idx.get(b, -1) for a, b in pairs]

    np.savez_compressed(
        os.path.join(OUT, "reference_c1.npz"),
        shape=np.array([n_users, n_items, d, k]), seed=np.array([100]),
        table_checksum=np.array([float(np.abs(table.astype(np.float64)).sum())]),
        table_rows=table[[0, 1, 942, 943, 2624]],
        viol=np.array([float(x) for x in viol]), viol_perturbed=np.array([float(x) for x in viol_p]),
        user_anchors=user_anchors, item_anchors=item_anchors,
        items_for_user_idx=np.array(fi_idx), items_for_user_score=np.array(fi_sc),
        similar_items_idx=np.array(fs_idx), similar_items_score=np.array(fs_sc),
        users_k200_idx=np.array(fu_idx), users_k200_score=np.array(fu_sc),
        posneg_idx=np.array(pn_idx), posneg_score=np.array(pn_sc),
        gcn_items_for_user_idx=np.array(gi_idx), gcn_items_for_user_score=np.array(gi_sc),
        pair_src=np.array(pair_src), pair_dst=np.array(pair_dst), pair_pred=pred.astype(np.float64),
    )

    # ------------------------------------------------------------------ case 2: metrics + extraction_efficiency
    y_true = {"a": 1, "b": 1, "c": 1}
    spot = np.array([ut.ndcg(y_true, ["x", "a", "b"]), ut.recall(y_true, ["x", "a", "b"]),
                     ut.reciprocal_rank(["a"], ["x", "a"]), ut.binary_ndcg({"a": 5.0, "b": 2.0}, ["b", "q", "a"]),
                     ut.ndcg({"s1": 5.0, "s2": 4.8, "s3": 3.0, "s4": 4.1, "s5": 2.9, "s6": 0.9},
                             ["s1", "s2", "s3", "s5", "s6"])])
    nu, ni, dd = 300, 500, 32
    _, collab2 = synthetic_case(nu, ni, dd, seed=200)
    table2 = ut.unit_length(collab2, axis=1)
    users2 = [Node("user", i) for i in range(nu)]
    items2 = [Node("item", i) for i in range(ni)]
    m = Dummy({"user", "item"}, n_dims=dd)
    m.add_nodes(users2 + items2)
    m.__build_knn__(table2)
    m.fit_done = True
    tr, vl = synthetic_edges(nu, ni, seed=300)
    train_edges = [Edge(users2[u], items2[i], w) for u, i, w in tr]
    val_edges = [Edge(users2[u], items2[i], w) for u, i, w in vl]
    random.seed(0)
    res = va.extraction_efficiency(m, train_edges, val_edges, va.model_get_topk, "item")
    met = res["metrics"]
    keys = ["recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"]
    # a few users' final filtered top-100 lists, to pin the filtering rule itself
    sample_users = [0, 1, 5, 35, 299]
    sample_preds = np.full((len(sample_users), 100), -1, dtype=np.int64)
    for j, u in enumerate(sample_users):
        p = res["predictions"].get(users2[u], [])
        sample_preds[j, :len(p)] = [int(x.node_external_id) for x in p]
    np.savez_compressed(
        os.path.join(OUT, "reference_eval.npz"),
        spot=spot, shape=np.array([nu, ni, dd]), seeds=np.array([200, 300]),
        metric_keys=np.array(keys), metric_values=np.array([float(met[x]) for x in keys]),
        sample_users=np.array(sample_users), sample_preds=sample_preds,
    )
    print("wrote", sorted(os.listdir(OUT)))
    print("metrics", {x: float(met[x]) for x in keys})
    print("spot", spot)


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/reference_c2.npz and reference_c3.npz by executing the
UNMODIFIED reference (loaded through oracle/ref_shim.py) at the shapes of BASELINE.json configs[1] and configs[2]:

  C2  ML-1M shape, 6,040 users x 3,706 items, d = 128: a full validation.extraction_efficiency run
      (hwer/validation.py:100-187: top-200 for every edge source, train items filtered, Recall@K / NDCG / diversity)
  C3  ML-20M item side, 27,278 items x d = 256 (+ 2,000 of the 138,493 users): find_closest_neighbours top-100
      for 96 user anchors and 16 item anchors (hwer/recommendation_base.py:157-174)

Run in the build container (the only place /root/reference exists):   python oracle/make_golden_c2c3.py
"""
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import synthetic_case, synthetic_edges  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
EE_KEYS = ["recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"]


def main():
    ref = ref_shim.load_reference()
    rb, ut, va = ref.recommendation_base, ref.utils, ref.validation
    Node, Edge = rb.Node, rb.Edge

    class Dummy(rb.RecommendationBase):
        def fit(self, *a, **k):
            pass

    # ------------------------------------------------------------------ C2
    nu, ni, d = 6040, 3706, 128
    _, collab = synthetic_case(nu, ni, d, seed=400)
    table = ut.unit_length(collab, axis=1)
    users = [Node("user", i) for i in range(nu)]
    items = [Node("item", i) for i in range(ni)]
    m = Dummy({"user", "item"}, n_dims=d)
    m.add_nodes(users + items)
    m.__build_knn__(table)
    m.fit_done = True
    tr, vl = synthetic_edges(nu, ni, seed=500)
    train = [Edge(users[u], items[i], w) for u, i, w in tr]
    val = [Edge(users[u], items[i], w) for u, i, w in vl]
    random.seed(0)
    t0 = time.time()
    res = va.extraction_efficiency(m, train, val, va.model_get_topk, "item")
    print("C2 reference extraction_efficiency: %.1f s, retrieval_time %.1f s" % (time.time() - t0, res["metrics"]["retrieval_time"]))
    met = res["metrics"]
    sample_users = [0, 7, 100, 3019, 6039]
    sample_preds = np.full((len(sample_users), 100), -1, dtype=np.int64)
    for j, u in enumerate(sample_users):
        p = res["predictions"].get(users[u], [])
        sample_preds[j, :len(p)] = [int(x.node_external_id) for x in p]
    np.savez_compressed(os.path.join(OUT, "reference_c2.npz"), shape=np.array([nu, ni, d]), seeds=np.array([400, 500]),
                        metric_keys=np.array(EE_KEYS), metric_values=np.array([float(met[k]) for k in EE_KEYS]),
                        retrieval_time=np.array([float(met["retrieval_time"])]),
                        sample_users=np.array(sample_users), sample_preds=sample_preds)
    print("C2 metrics", {k: float(met[k]) for k in EE_KEYS})

    # ------------------------------------------------------------------ C3 (item side in full, 2,000 users)
    nu3, ni3, d3, k3 = 2000, 27278, 256, 100
    _, collab3 = synthetic_case(nu3, ni3, d3, seed=600)
    table3 = ut.unit_length(collab3, axis=1)
    users3 = [Node("user", i) for i in range(nu3)]
    items3 = [Node("item", i) for i in range(ni3)]
    m3 = Dummy({"user", "item"}, n_dims=d3)
    m3.add_nodes(users3 + items3)
    m3.__build_knn__(table3)
    m3.fit_done = True
    rs = np.random.RandomState(9)
    ua = rs.choice(nu3, 96, replace=False)
    ia = rs.choice(ni3, 16, replace=False)

    def run(anchor):
        r = m3.find_closest_neighbours("item", anchor, k=k3)
        return [int(n.node_external_id) for n, s in r], [float(s) for n, s in r]

    t0 = time.time()
    u_idx, u_sc = zip(*[run(users3[u]) for u in ua])
    i_idx, i_sc = zip(*[run(items3[i]) for i in ia])
    print("C3 reference: %.1f ms per find_closest_neighbours" % ((time.time() - t0) * 1e3 / (len(ua) + len(ia))))
    np.savez_compressed(os.path.join(OUT, "reference_c3.npz"), shape=np.array([nu3, ni3, d3, k3]), seed=np.array([600]),
                        user_anchors=ua, item_anchors=ia, user_idx=np.array(u_idx), user_score=np.array(u_sc),
                        item_idx=np.array(i_idx), item_score=np.array(i_sc))


if __name__ == "__main__":
    main()

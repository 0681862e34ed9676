"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/reference_lp.npz by executing the UNMODIFIED reference's
`validation.link_prediction_accuracy` and `validation.get_prediction_details` (hwer/validation.py:41-65,258-275,
loaded through oracle/ref_shim.py) on the seeded synthetic graph of make_golden.py's case 2.

Run in the build container (the only place /root/reference exists):   python oracle/make_golden_lp.py
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import synthetic_case, synthetic_edges  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
LP_KEYS = ["lp_train_ap", "lp_val_ap", "lp_train_precision", "lp_train_recall", "lp_val_precision", "lp_val_recall",
           "lp_train_accuracy", "lp_val_accuracy"]
EE_KEYS = ["recall@100", "ndcg_b@100", "ndcg_b@10", "recall@10", "diversity"]


def main():
    ref = ref_shim.load_reference()
    rb, ut, va = ref.recommendation_base, ref.utils, ref.validation
    Node, Edge = rb.Node, rb.Edge

    class Dummy(rb.RecommendationBase):
        def fit(self, *a, **k):
            pass

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
    nodes = users2 + items2

    random.seed(11)
    lp = va.link_prediction_accuracy(m, nodes, train_edges, val_edges)          # validation.py:41-65
    # a second draw with a different seed, and a degenerate score set with many ties (a table of few distinct rows)
    random.seed(12)
    lp2 = va.link_prediction_accuracy(m, nodes, train_edges[:500], val_edges[:50])

    random.seed(13)
    preds, actuals, stats = va.get_prediction_details(m, nodes, train_edges, val_edges, va.model_get_topk, "item")

    np.savez_compressed(
        os.path.join(OUT, "reference_lp.npz"),
        shape=np.array([nu, ni, dd]), seeds=np.array([200, 300, 11, 12, 13]),
        lp_keys=np.array(LP_KEYS), lp_values=np.array([float(lp[k]) for k in LP_KEYS]),
        lp2_values=np.array([float(lp2[k]) for k in LP_KEYS]),
        details_predictions=np.asarray(preds, dtype=np.float64), details_actuals=np.asarray(actuals, dtype=np.float64),
        details_lp_values=np.array([float(stats[k]) for k in LP_KEYS]),
        ee_keys=np.array(EE_KEYS), details_ee_values=np.array([float(stats[k]) for k in EE_KEYS]),
        details_keys=np.array(sorted(stats.keys())),
    )
    print("lp", {k: float(lp[k]) for k in LP_KEYS})
    print("lp2", {k: float(lp2[k]) for k in LP_KEYS})
    print("details keys", sorted(stats.keys()))


if __name__ == "__main__":
    main()

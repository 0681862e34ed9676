"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the NCF re-rank row, produced by running the
UNMODIFIED reference in this container: hwer/ncf.py NCF.forward, and GcnNCF.predict / find_closest_neighbours with
ncf_enabled (hwer/gcn_ncf.py:329-361,363-387), through oracle/ref_shim.py.

    python oracle/make_golden_ncf.py        ->  tests/golden/reference_ncf.npz
"""
import importlib
import os

import numpy as np
import torch

import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def flat_params(model):
    """[W1, b1, ..., W_depth, b_depth, w_out, b_out] in the layout of include/hwer_b200.h (torch Linear: out x in)."""
    lin = [m for m in model.W if isinstance(m, torch.nn.Linear)]
    return np.concatenate([np.concatenate([l.weight.detach().numpy().reshape(-1), l.bias.detach().numpy().reshape(-1)])
                           for l in lin]).astype(np.float32)


def main():
    ref = ref_shim.load_reference()
    rb, gn = ref.recommendation_base, ref.gcn_ncf
    ncf_mod = importlib.import_module("hwer.ncf")
    Node = rb.Node
    out = {}
    F = 32
    n_users, n_items = 40, 60
    rs = np.random.RandomState(11)
    # the NCF's input table: prediction_artifacts["h"], row 0 = padding node (gcn_ncf.py:227,339-341)
    h = rs.standard_normal((n_users + n_items + 1, F)).astype(np.float32) * 0.5
    out["h"] = h
    src = rs.randint(0, n_users + n_items + 1, 300)
    dst = rs.randint(0, n_users + n_items + 1, 300)
    src[:5] = 0
    dst[3:8] = 0
    out["src"], out["dst"] = src.astype(np.int64), dst.astype(np.int64)
    for depth in (1, 2, 3, 4):
        torch.manual_seed(100 + depth)
        model = ncf_mod.NCF(F, depth, 0.1)                 # ncf.py:8-22, the reference's own initialisation
        with torch.no_grad():
            for m in model.W:
                if isinstance(m, torch.nn.Linear):
                    m.bias.add_(torch.randn(m.bias.shape) * 0.1)      # biases start at zero: make them count
        model.eval()
        with torch.no_grad():
            y = model(torch.tensor(src), torch.tensor(dst), torch.tensor(h[src]), torch.tensor(h[dst])).numpy()
        out["params_d%d" % depth] = flat_params(model)
        out["forward_d%d" % depth] = y.astype(np.float64)
        if depth == 3:
            model3 = model
    # GcnNCF serving with the NCF branch (depth 3)
    users = [Node("user", i) for i in range(n_users)]
    items = [Node("item", i) for i in range(n_items)]
    g = gn.GcnNCF({}, {"user", "item"}, n_dims=F)
    g.add_nodes(users + items)
    table = h[1:] / np.linalg.norm(h[1:], axis=1, keepdims=True)      # knn vectors: unit rows of the same table
    g.__build_knn__(table.astype(np.float32))
    g.fit_done = True
    g.ncf_enabled = True
    g.prediction_artifacts = {"model": model3, "h": torch.tensor(h)}
    pairs = [(users[int(a)], items[int(b)]) for a, b in zip(rs.randint(0, n_users, 50), rs.randint(0, n_items, 50))]
    pairs += [(Node("user", "ghost"), items[5]), (users[4], Node("item", "ghost"))]
    out["predict"] = np.asarray(g.predict(pairs), dtype=np.float64)                 # gcn_ncf.py:336-361
    out["predict_src"] = np.array([g.nodes_to_idx.get(a, -1) for a, b in pairs])
    out["predict_dst"] = np.array([g.nodes_to_idx.get(b, -1) for a, b in pairs])
    anchors = [0, 7, 19, 39]
    k = 20
    idx, sc = [], []
    for u in anchors:
        res = g.find_closest_neighbours("item", users[u], k=k)                       # gcn_ncf.py:363-387, NCF branch
        idx.append([int(n.node_external_id) for n, s in res])
        sc.append([float(s) for n, s in res])
    out["table"] = table.astype(np.float32)
    out["anchors"], out["fcn_idx"], out["fcn_score"] = np.array(anchors), np.array(idx), np.array(sc)
    out["shape"] = np.array([n_users, n_items, F, k])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "reference_ncf.npz"), **out)
    print("wrote reference_ncf.npz", {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()

"""Host-side mirror of the hot-path callers in hwer/validation.py, batched onto the GPU.

  model_get_topk_gpu        replaces model_get_topk_knn          hwer/validation.py:30-35 (hook alias :38)
  extraction_efficiency     same signature and `metrics` keys    hwer/validation.py:100-187
  ncf_eval                  1 positive + 100 sampled negatives   hwer/validation.py:68-97
The per-user Python loops over reciprocal_rank / ndcg / recall (hwer/utils.py:71-121) become one
hwer_eval_metrics launch; pair scores come from hwer_pair_score.
"""
import random
import time
from collections import defaultdict
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops
from .recommendation_base import Edge, Node, NodeType, RecommendationBase


def model_get_topk_gpu(model: RecommendationBase, anchors: List[Node], node_type: NodeType, k: int = 200
                       ) -> Dict[Node, List[Tuple[Node, float]]]:
    predictions = defaultdict(list)
    if len(anchors) == 0:
        return predictions
    rows, scores = model.find_closest_neighbours_batch(node_type, list(anchors), k=k)
    for u, p in zip(anchors, model.rows_to_nodes(rows, scores)):
        predictions[u] = p
    return predictions


model_get_topk = model_get_topk_gpu


def _csr(lists, device):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    flat = np.fromiter((x for l in lists for x in l), dtype=np.int64, count=int(ptr[-1]))
    return torch.from_numpy(ptr).to(device), torch.from_numpy(flat).to(device)


def ranking_metrics(model: RecommendationBase, users: List[Node], topk_rows: torch.Tensor, train_edges, validation_edges,
                    node_type: NodeType, cutoffs=(10, 20, 50, 100)):
    """Device evaluation of `topk_rows` ([U, k] global rows in rank order for `users`).  Returns a dict with
    recall@c / ndcg@c / ndcg_b@c for every cutoff, mrr, diversity and the number of validation users."""
    dev = topk_rows.device
    local = {int(g): i for i, g in enumerate(model.knn.idxs[node_type])}     # global row -> item id within type
    uidx = {u: i for i, u in enumerate(users)}
    n2i = model.nodes_to_idx
    train = [set() for _ in users]
    for u, i, r in train_edges:
        if u in uidx and i in n2i and n2i[i] in local:
            train[uidx[u]].add(local[n2i[i]])
    val = [dict() for _ in users]
    for u, i, r in validation_edges:
        if u in uidx:
            # an item outside the index can never be retrieved but still counts as a true item: id past the table
            item = local[n2i[i]] if (i in n2i and n2i[i] in local) else len(local) + len(val[uidx[u]])
            val[uidx[u]][item] = float(r)     # dict semantics of validation.py:160 (last rating wins)
    val_sorted = [sorted(v.items(), key=lambda x: -x[1]) for v in val]
    train_ptr, train_idx = _csr([sorted(t) for t in train], dev)
    val_ptr, val_idx = _csr([[i for i, r in v] for v in val_sorted], dev)
    val_rel = torch.tensor([r for v in val_sorted for i, r in v], dtype=torch.float32, device=dev)
    # global rows -> ids within the type
    off = model.knn.offset[node_type]
    if off is not None:
        items = torch.where(topk_rows >= 0, topk_rows - off, topk_rows)
    else:
        lut = torch.full((len(n2i),), -1, dtype=torch.int64, device=dev)
        lut[model.knn.idxs_dev[node_type]] = torch.arange(len(local), device=dev)
        items = torch.where(topk_rows >= 0, lut[topk_rows.clamp(min=0)], topk_rows)
    cutoffs = sorted(cutoffs)
    out = ops.eval_metrics(items.contiguous(), train_ptr, train_idx, val_ptr, val_idx, val_rel, cutoffs,
                           len(local)).cpu().tolist()
    res = {}
    for j, c in enumerate(cutoffs):
        res["recall@%d" % c], res["ndcg@%d" % c], res["ndcg_b@%d" % c] = out[3 * j], out[3 * j + 1], out[3 * j + 2]
    res["mrr"] = out[3 * len(cutoffs)]
    res["distinct_items"] = out[3 * len(cutoffs) + 1]
    res["validation_users"] = out[3 * len(cutoffs) + 2]
    return res


def ncf_eval(model: RecommendationBase, train_edges: List[Edge], validation_edges: List[Edge], item_list: List[Node]):
    item_list = set(item_list)
    interactions = defaultdict(set)
    for u, i, _ in train_edges:
        interactions[u].add(i)
    for u, i, _ in validation_edges:
        interactions[u].add(i)
    user_test_item = {}
    for u, i, _ in validation_edges:     # one entry per user, the last validation edge wins (validation.py:79-81)
        pool = sorted(item_list - interactions[u], key=repr)
        user_test_item[u] = [i, *random.sample(pool, 100)]
    if not user_test_item:
        return {"ncf_hr": float("nan"), "ncf_ndcg": float("nan")}
    users = list(user_test_item.keys())
    src = model._rows_of([u for u in users for _ in range(101)])
    dst = model._rows_of([it for u in users for it in user_test_item[u]])
    s = model.predict_rows(src, dst).reshape(len(users), 101)
    # stable descending sort keeps the positive (column 0) ahead of equal-scored negatives (validation.py:85)
    rank = (s[:, 1:] > s[:, :1]).sum(dim=1)
    hit = rank < 10
    ndcg = torch.where(hit, 1.0 / torch.log2(rank.double() + 2.0) / (1.0 + 1e-8), torch.zeros_like(rank, dtype=torch.float64))
    return {"ncf_hr": float(hit.double().mean().item()), "ncf_ndcg": float(ndcg.mean().item())}


def extraction_efficiency(model, train_edges: List[Edge], validation_edges: List[Edge], get_topk=None,
                          node_type: NodeType = "item", k: int = 200):
    train_users = list(set([u for u, i, r in train_edges]))
    validation_users = list(set([u for u, i, r in validation_edges]))
    all_users = list(set(train_users + validation_users))
    all_items = list(set([i for u, i, r in validation_edges] + [i for u, i, r in train_edges]))
    all_items = [x for x in all_items if x.node_type == node_type]
    s = time.time()
    rows, scores = model.find_closest_neighbours_batch(node_type, all_users, k=k)
    torch.cuda.synchronize()
    pred_time = time.time() - s
    m = ranking_metrics(model, all_users, rows, train_edges, validation_edges, node_type)
    ncf_metrics = ncf_eval(model, train_edges, validation_edges, all_items)
    metrics = {"retrieval_time": pred_time,
               "recall@100": m["recall@100"],
               "ndcg_b@100": m["ndcg_b@100"],
               "ndcg_b@10": m["ndcg_b@10"],
               "recall@10": m["recall@10"],
               "diversity": m["distinct_items"] / max(len(all_items), 1), **ncf_metrics}
    return {"users": all_users, "rows": rows, "scores": scores, "metrics": metrics, "all_metrics": m}

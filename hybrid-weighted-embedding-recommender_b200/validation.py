"""Host-side mirror of the hot-path callers in hwer/validation.py, batched onto the GPU.

  model_get_topk_gpu        replaces model_get_topk_knn          hwer/validation.py:30-35 (hook alias :38)
  extraction_efficiency     same signature and `metrics` keys    hwer/validation.py:100-187
  ncf_eval                  1 positive + 100 sampled negatives   hwer/validation.py:68-97
  link_prediction_accuracy  10x sampled negative pairs           hwer/validation.py:41-65
  get_prediction_details, test_algorithm, test_multiple_algorithms, display_results, run_model_for_hpo,
  run_models_for_testing    the harness around them              hwer/validation.py:190-309
The per-user Python loops over reciprocal_rank / ndcg / recall (hwer/utils.py:71-121) become one
hwer_eval_metrics launch; pair scores come from hwer_pair_score (or hwer_ncf_score); the sklearn metric calls of
link_prediction_accuracy become one hwer_link_metrics launch per pair set.
"""
import copy
import datetime
import random
import time
from collections import defaultdict
from typing import Any, Dict, List, Set, Tuple

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


def _edge_arrays(model, edges):
    """(src row, dst row, weight) numpy arrays of an edge list (-1 = node unknown to the model) and the dst nodes.
    The one per-edge Python pass of the evaluation; everything after it is array work."""
    n2i = model.nodes_to_idx
    m = len(edges)
    src = np.empty(m, dtype=np.int64)
    dst = np.empty(m, dtype=np.int64)
    w = np.empty(m, dtype=np.float64)
    dst_nodes = [None] * m
    for j, (u, i, r) in enumerate(edges):
        src[j] = n2i.get(u, -1)
        dst[j] = n2i.get(i, -1)
        w[j] = r
        dst_nodes[j] = i
    return src, dst, w, dst_nodes


def _csr_from_sorted(owner, n_owners):
    """ptr [n_owners + 1] of entries already grouped by ascending `owner`."""
    counts = np.bincount(owner, minlength=n_owners) if owner.size else np.zeros(n_owners, dtype=np.int64)
    ptr = np.zeros(n_owners + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr


def ranking_metrics(model: RecommendationBase, users: List[Node], topk_rows: torch.Tensor, train_edges, validation_edges,
                    node_type: NodeType, cutoffs=(10, 20, 50, 100)):
    """Device evaluation of `topk_rows` ([U, k] global rows in rank order for `users`).  Returns a dict with
    recall@c / ndcg@c / ndcg_b@c for every cutoff, mrr, diversity and the number of validation users.
    Host side: one pass over the edges to look their nodes up, then numpy builds the two CSR structures
    hwer_eval_metrics reads (train items per user ascending; validation items per user by relevance descending,
    the dict semantics of validation.py:158-161: per (user, item) the last rating wins)."""
    dev = topk_rows.device
    n_rows = len(model.nodes_to_idx)
    type_rows = model.knn.idxs[node_type]
    n_local = int(type_rows.shape[0])
    local_of = np.full(n_rows + 1, -1, dtype=np.int64)         # global row -> id within the type; [-1] stays -1
    local_of[type_rows] = np.arange(n_local, dtype=np.int64)
    user_rows = np.fromiter((model.nodes_to_idx[u] for u in users), dtype=np.int64, count=len(users))
    upos = np.full(n_rows + 1, -1, dtype=np.int64)             # global row -> position in `users`
    upos[user_rows] = np.arange(len(users), dtype=np.int64)
    U = len(users)

    # training items per user: unique (user, item) pairs, ascending
    t_src, t_dst, _, _ = _edge_arrays(model, train_edges)
    tu, ti = upos[t_src], local_of[t_dst]
    ok = (tu >= 0) & (ti >= 0)
    tkey = np.unique(tu[ok] * max(n_local, 1) + ti[ok])
    train_ptr = _csr_from_sorted(tkey // max(n_local, 1), U)
    train_idx = tkey % max(n_local, 1)

    # validation items per user
    v_src, v_dst, v_w, v_nodes = _edge_arrays(model, validation_edges)
    vu, vi = upos[v_src], local_of[v_dst]
    # an item outside the index can never be retrieved but still counts as a true item: ids past the table, one
    # per distinct node
    extra = {}
    for j in np.flatnonzero((vu >= 0) & (vi < 0)):
        vi[j] = n_local + extra.setdefault(v_nodes[j], len(extra))
    ok = vu >= 0
    vu, vi, vw = vu[ok], vi[ok], v_w[ok]
    span = n_local + len(extra) + 1
    vkey = vu * span + vi
    uniq, first = np.unique(vkey, return_index=True)                      # first time the pair appears
    _, last_rev = np.unique(vkey[::-1], return_index=True)                # last time: its rating wins
    rel = vw[vkey.shape[0] - 1 - last_rev] if vkey.size else vw
    order = np.lexsort((first, -rel, uniq // span))                       # by user, relevance desc, insertion order
    val_ptr = _csr_from_sorted((uniq // span)[order], U)
    val_idx = (uniq % span)[order]
    val_rel = rel[order].astype(np.float32)

    # global rows -> ids within the type
    off = model.knn.offset[node_type]
    if off is not None:
        items = ops.map_rows(topk_rows.contiguous(), None, -off)
    else:
        lut = torch.from_numpy(local_of[:-1].copy()).to(dev)
        items = ops.map_rows(topk_rows.contiguous(), lut)
    cutoffs = sorted(cutoffs)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = ops.eval_metrics(items, to_dev(train_ptr), to_dev(train_idx), to_dev(val_ptr), to_dev(val_idx),
                           to_dev(val_rel), cutoffs, max(n_local, 1)).cpu().tolist()
    res = {}
    for j, c in enumerate(cutoffs):
        res["recall@%d" % c], res["ndcg@%d" % c], res["ndcg_b@%d" % c] = out[3 * j], out[3 * j + 1], out[3 * j + 2]
    res["mrr"] = out[3 * len(cutoffs)]
    res["distinct_items"] = out[3 * len(cutoffs) + 1]
    res["validation_users"] = out[3 * len(cutoffs) + 2]
    return res


def _negative_pool(ordered_items, seen):
    """Items the user has not interacted with, in the order of `ordered_items` -- the same list as
    sorted(set(items) - seen, key=repr) when `ordered_items` is sorted by repr, without sorting per user."""
    return [x for x in ordered_items if x not in seen]


def _sample_negatives(validation_edges, interactions, ordered_items):
    """For every validation edge, 100 items the edge's user never touched, drawn like
    random.sample(sorted(items - seen, key=repr), 100) (validation.py:80 under Python >= 3.11, where sampling
    from a set needs an order): same draws from the `random` module, same picks, but the pool is a numpy index
    vector -- random.sample(range(n), k) returns the positions random.sample(pool, k) would read.  Returns
    {user: (positive item, [100 positions into ordered_items])}; one entry per user, the last edge wins (:79-81)."""
    n_items = len(ordered_items)
    position = {it: j for j, it in enumerate(ordered_items)}
    seen_positions = {}
    out = {}
    for u, i, _ in validation_edges:
        sp = seen_positions.get(u)
        if sp is None:
            sp = np.fromiter((position[x] for x in interactions[u] if x in position), dtype=np.int64)
            seen_positions[u] = sp
        free = np.ones(n_items, dtype=bool)
        free[sp] = False
        pool = np.flatnonzero(free)
        out[u] = (i, pool[random.sample(range(pool.shape[0]), 100)])
    return out


def ncf_eval_scores(model: RecommendationBase, train_edges: List[Edge], validation_edges: List[Edge],
                    item_list: List[Node]):
    """The scored candidate lists of ncf_eval (hwer/validation.py:68-84): [U, 101] scores, column 0 = the user's
    positive, columns 1..100 = the sampled negatives, all U x 101 pairs scored in one device call.
    None if there is no validation edge."""
    interactions = defaultdict(set)
    for u, i, _ in train_edges:
        interactions[u].add(i)
    for u, i, _ in validation_edges:
        interactions[u].add(i)
    ordered_items = sorted(set(item_list), key=repr)
    picks = _sample_negatives(validation_edges, interactions, ordered_items)
    if not picks:
        return None
    users = list(picks.keys())
    item_rows = model._rows_of(ordered_items)                                   # rows of the candidate items, once
    user_rows = model._rows_of(users)
    pos_rows = model._rows_of([picks[u][0] for u in users])
    neg_pos = torch.from_numpy(np.stack([picks[u][1] for u in users])).to(item_rows.device)     # [U, 100]
    dst = torch.cat([pos_rows[:, None], item_rows[neg_pos]], dim=1).reshape(-1).contiguous()   # positive first
    src = user_rows[:, None].expand(len(users), 101).reshape(-1).contiguous()
    return model.predict_rows(src, dst).float().reshape(len(users), 101).contiguous()


def ncf_eval(model: RecommendationBase, train_edges: List[Edge], validation_edges: List[Edge], item_list: List[Node]):
    """hwer/validation.py:68-97: per validation user 1 positive + 100 sampled negatives, scored by the model,
    HR@10 and binary NDCG@10 of the positive's rank.  The rank of the positive is the number of negatives scored
    strictly higher (the reference's stable descending sort keeps the positive, listed first, ahead of equal
    scores, :82-85); ranks, HR@10 and NDCG@10 are computed by hwer_hit_rank_metrics on the device."""
    s = ncf_eval_scores(model, train_edges, validation_edges, item_list)
    if s is None:
        return {"ncf_hr": float("nan"), "ncf_ndcg": float("nan")}
    hr, ndcg = ops.hit_rank_metrics(s, topn=10).tolist()
    return {"ncf_hr": hr, "ncf_ndcg": ndcg}


def _rows_from_predictions(model, users, predictions, device):
    """Dict[Node, List[(Node, score)]] (what a `get_topk` hook returns, validation.py:30-35) -> [U, kmax] global rows
    in descending score order (stable, validation.py:134-135), padded with -1."""
    n2i = model.nodes_to_idx
    lists = []
    for u in users:
        p = sorted(predictions.get(u, []), key=lambda x: x[1], reverse=True)
        lists.append([n2i.get(n, -1) for n, _ in p])
    kmax = max([len(l) for l in lists] + [1])
    rows = np.full((len(users), kmax), -1, dtype=np.int64)
    for j, l in enumerate(lists):
        rows[j, :len(l)] = l
    return torch.from_numpy(rows).to(device)


def extraction_efficiency(model, train_edges: List[Edge], validation_edges: List[Edge], get_topk=None,
                          node_type: NodeType = "item", k: int = 200):
    """hwer/validation.py:100-187.  `get_topk(model, anchors, node_type) -> Dict[Node, List[(Node, score)]]` is the
    reference's retrieval hook (:100,111).  None or this module's own `model_get_topk` take the tensor path (one
    batched search, no Python tuples); any other callable is called exactly as the reference calls it and its
    lists are evaluated."""
    train_users = list(set([u for u, i, r in train_edges]))
    validation_users = list(set([u for u, i, r in validation_edges]))
    all_users = list(set(train_users + validation_users))
    all_items = list(set([i for u, i, r in validation_edges] + [i for u, i, r in train_edges]))
    all_items = [x for x in all_items if x.node_type == node_type]
    s = time.time()
    if get_topk is None or get_topk is model_get_topk_gpu:
        rows, scores = model.find_closest_neighbours_batch(node_type, all_users, k=k)
        torch.cuda.synchronize()
    else:
        predictions = get_topk(model, all_users, node_type)
        rows, scores = _rows_from_predictions(model, all_users, predictions, model.device_vectors.device), None
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


def link_prediction_accuracy(model: RecommendationBase, nodes: List[Node], train_edges: List[Edge],
                             validation_edges: List[Edge]):
    """hwer/validation.py:41-65.  The negative pairs are drawn on the host with random.choices in the reference's
    order (train src, train dst, validation src, validation dst), so `random.seed` reproduces its pair sets; the
    pair scores and every metric are computed on the device."""
    m = 10
    sets = []
    for edges in (train_edges, validation_edges):
        neg_src = random.choices(nodes, k=len(edges) * m)
        neg_dst = random.choices(nodes, k=len(edges) * m)
        sets.append(([e.src for e in edges] + neg_src, [e.dst for e in edges] + neg_dst, len(edges)))
    results = {}
    for name, (src, dst, n_pos) in zip(("train", "val"), sets):
        if len(src) == 0:
            raise ValueError("link_prediction_accuracy: empty %s edge list" % name)
        scores = model.predict_rows(model._rows_of(src), model._rows_of(dst))
        labels = torch.zeros(len(src), dtype=torch.uint8, device=scores.device)
        labels[:n_pos] = 1
        ap, precision, recall, accuracy = ops.link_metrics(scores.float().contiguous(), labels)[:4].tolist()
        results["lp_%s_ap" % name] = ap
        results["lp_%s_precision" % name] = precision
        results["lp_%s_recall" % name] = recall
        results["lp_%s_accuracy" % name] = accuracy
    return results


def _pair_predictions(recsys, affinities):
    """Scores of the (src, dst) pairs of `affinities` and their weights; NaN scores are an error (:259-267)."""
    scores = np.asarray(recsys.predict([(e[0], e[1]) for e in map(tuple, affinities)]))
    bad = np.isnan(scores)
    if bad.any():
        raise AssertionError("Encountered Nan Predictions = %s" % int(bad.sum()),
                             [tuple(e) for e, b in zip(affinities, bad) if b])
    return scores, np.array([tuple(e)[2] for e in affinities])


def get_prediction_details(recsys, nodes: List[Node], train_affinities: List[Edge], validation_affinities: List[Edge],
                           model_get_topk=None, node_type: NodeType = "item"):
    """hwer/validation.py:258-275: pair scores of the validation edges, their weights, and one metric dict holding
    the link-prediction metrics and extraction_efficiency's `metrics`."""
    predictions, actuals = _pair_predictions(recsys, validation_affinities)
    _pair_predictions(recsys, train_affinities)                 # the reference scores them too (NaN check only)
    retrieval = extraction_efficiency(recsys, train_affinities, validation_affinities, model_get_topk, node_type)
    stats = link_prediction_accuracy(recsys, nodes, train_affinities, validation_affinities)
    stats.update(retrieval["metrics"])
    return predictions, actuals, stats


_PROBE_IDS = ("eifjcchchbniufclvfdugvhnftdvjculhjitjihuncce", "eifjcchchbnirdjknkrvtfkbfurvjdfjhllbddtbvicb")


def test_algorithm(train_affinities: List[Edge], validation_affinities: List[Edge],
                   nodes: List[Node], node_types: Set[NodeType], hyperparameters,
                   get_data_mappers, algo, node_type: NodeType):
    """hwer/validation.py:190-222.  Training is outside this package: `hyperparameters` carries the trained tables
    (`vectors` for algo "content"; `collaborative_vectors` [+ `content_vectors`, `alpha`] for "gcn_ncf") next to
    `n_dims`; fit() picks them up from its `hyperparameters` keyword."""
    from . import ContentRecommendation, GcnNCF
    embedding_mapper, node_data = get_data_mappers()
    cls = {"gcn_ncf": GcnNCF, "content": ContentRecommendation}[algo]
    recsys = cls(embedding_mapper=embedding_mapper, node_types=node_types, n_dims=hyperparameters["n_dims"])
    t0 = time.time()
    recsys.fit(nodes, train_affinities, node_data, hyperparameters=copy.copy(hyperparameters))
    fit_seconds = time.time() - t0

    # pairs with nodes that were never trained on must still score (to ~0.5), never NaN (:204-213)
    some_type = list(node_types)[0]
    ghost_a, ghost_b = (Node(some_type, i) for i in _PROBE_IDS)
    first = train_affinities[0]
    default_preds = recsys.predict([(first.src, ghost_a), (first.src, first.dst), (ghost_a, ghost_b), (ghost_b, first.src)])
    print("Default Preds = ", default_preds)
    assert not np.isnan(np.asarray(default_preds)).any()

    predictions, actuals, stats = get_prediction_details(recsys, nodes, train_affinities, validation_affinities,
                                                         model_get_topk, node_type)
    return recsys, [dict({"algo": algo, "time": fit_seconds}, **stats)], predictions, actuals


test_algorithm.__test__ = False      # a harness entry point named like the reference's, not a pytest test


def test_multiple_algorithms(train_affinities, validation_affinities, nodes: List[Node], node_types: Set[NodeType],
                             hyperparamters_dict, get_data_mappers, algos, node_type: NodeType):
    """hwer/validation.py:225-240."""
    wanted = set(algos)
    assert wanted and wanted <= {"content", "gcn_ncf"}
    recs, results = [], []
    for algo in wanted:
        rec, res, _, _ = test_algorithm(train_affinities, validation_affinities, nodes, node_types,
                                        hyperparamters_dict[algo], get_data_mappers, algo, node_type)
        recs.append(rec)
        results += res
    return recs, results


test_multiple_algorithms.__test__ = False


def _as_duration(seconds):
    return str(datetime.timedelta(seconds=seconds))


def display_results(results: List[Dict[str, Any]]):
    """hwer/validation.py:243-255: per-algorithm means, printed eight columns at a time; returns the frame with
    `retrieval_time` left numeric."""
    import pandas as pd
    from tabulate import tabulate
    frame = pd.DataFrame.from_records(results).groupby("algo").mean()
    seconds = frame["retrieval_time"].copy()
    frame["time"] = frame["time"].map(_as_duration)
    frame["retrieval_time"] = seconds.map(_as_duration)
    names = list(frame.columns)
    for first in range(0, len(names), 8):
        print(tabulate(frame[names[first:first + 8]], headers="keys", tablefmt="psql"))
    frame["retrieval_time"] = seconds
    return frame


def run_model_for_hpo(nodes: List[Node], edges: List[Tuple[Edge, bool]], node_types: Set[NodeType],
                      retrieved_node_type: NodeType, prepare_data_mappers, hyperparameters, algo):
    """hwer/validation.py:278-287."""
    return run_models_for_testing(nodes, edges, node_types, retrieved_node_type, prepare_data_mappers, [algo],
                                  {algo: hyperparameters}, display=False)


def run_models_for_testing(nodes: List[Node], edges: List[Tuple[Edge, bool]], node_types: Set[NodeType],
                           retrieved_node_type: NodeType, prepare_data_mappers, algos, hyperparamters_dict,
                           display=True, results_csv="overall_results.csv"):
    """hwer/validation.py:290-309: `edges` are (Edge, is_validation) pairs; returns (ndcg_b@100, ncf_ndcg) of the
    first algorithm.  display=True prints the table and writes it to `results_csv` (None: skip the file)."""
    import pandas as pd
    train = [e for e, held_out in edges if not held_out]
    held = [e for e, held_out in edges if held_out]
    _, results = test_multiple_algorithms(train, held, nodes, node_types, hyperparamters_dict, prepare_data_mappers,
                                          algos, retrieved_node_type)
    if display:
        headline = results[0]["ndcg_b@100"], results[0]["ncf_ndcg"]
        table = display_results(results)
        if results_csv:
            table.to_csv(results_csv)
        return headline
    means = pd.DataFrame.from_records(results).groupby("algo").mean().reset_index()
    return means["ndcg_b@100"].values[0], means["ncf_ndcg"].values[0]

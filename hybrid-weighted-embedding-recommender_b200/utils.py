"""Host-side mirror of the hot-path helpers of hwer/utils.py, backed by the CUDA kernels.

Same names, argument meaning and return shapes as the reference so call sites read the same:
  unit_length(a, axis)                   hwer/utils.py:43-44 (+ repeat_args_wrapper :286-293,312)
  unit_length_violations(a, axis, eps)   hwer/utils.py:51-57
  NodeNotFoundException                  hwer/utils.py:326
  reciprocal_rank, average_precision, ndcg, binary_ndcg, binary_ndcg_v2, recall
                                         hwer/utils.py:71-121 -- the names hwer/validation.py:26 imports; one
                                         list at a time through the same hwer_eval_metrics kernel that
                                         validation.extraction_efficiency runs for all users at once
numpy inputs are uploaded to the current CUDA device and results come back as numpy; CUDA tensors stay on the
device.  Nothing here computes on the CPU.
"""
import numpy as np
import torch

from . import ops


class NodeNotFoundException(Exception):
    pass


def _to_device(a):
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.cuda()
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _unit_length_one(a, axis=0):
    was_np = not isinstance(a, torch.Tensor)
    src_dtype = a.dtype
    t = _to_device(a)
    squeeze = False
    if t.dim() == 1:
        t, squeeze = t[None, :], True
    elif axis == 0:
        t = t.t().contiguous()
    out = ops.unit_length(t)
    if squeeze:
        out = out[0]
    elif axis == 0:
        out = out.t().contiguous()
    if was_np:
        return out.cpu().numpy().astype(src_dtype, copy=False)
    return out


def unit_length(*args, **kwargs):
    results = [_unit_length_one(a, **kwargs) for a in args]
    return results[0] if len(results) == 1 else results


def _violations_one(a, axis=0, epsilon=1e-4):
    t = _to_device(a)
    if t.dim() == 1:
        t = t[None, :]
    elif axis == 0:
        t = t.t().contiguous()
    v, mean, pos, neg, _ = ops.norm_stats(t, epsilon)
    return v, mean, pos, neg


def unit_length_violations(*args, **kwargs):
    results = [_violations_one(a, **kwargs) for a in args]
    return results[0] if len(results) == 1 else results


# --------------------------------------------------------------------------- ranking metrics of ONE list
_MAX_LIST = 256      # hwer_eval_metrics' largest cutoff


def _one_list(y_true_rel, y_pred):
    """(recall, graded ndcg, binary ndcg, reciprocal rank) of one prediction list on the device.
    y_true_rel: {item: relevance}; y_pred: ranked items (hwer/utils.py:71-121 semantics: both ndcg variants
    truncate the ideal list to len(y_pred), recall divides by min(|pred|, |true|))."""
    y_pred = list(y_pred)
    if len(y_pred) == 0:
        return 0.0, 0.0, 0.0, 0.0
    if len(y_pred) > _MAX_LIST:
        raise ValueError("ranking metrics are evaluated on lists of at most %d predictions" % _MAX_LIST)
    true_sorted = sorted(y_true_rel.items(), key=lambda kv: -kv[1])
    ids = {k: j for j, (k, _) in enumerate(true_sorted)}
    T = len(true_sorted)
    extra = {}
    pred_ids = []
    for it in y_pred:
        if it in ids:
            pred_ids.append(ids[it])
        else:
            pred_ids.append(extra.setdefault(it, T + len(extra)))
    dev = torch.device("cuda", torch.cuda.current_device())
    i64 = dict(dtype=torch.int64, device=dev)
    topk = torch.tensor([pred_ids], **i64)
    zero_ptr = torch.zeros(2, **i64)
    val_ptr = torch.tensor([0, T], **i64)
    val_idx = torch.arange(max(T, 1), **i64)
    val_rel = torch.tensor([float(r) for _, r in true_sorted] or [0.0], dtype=torch.float32, device=dev)
    _, per_user = ops.eval_metrics(topk, zero_ptr, torch.zeros(1, **i64), val_ptr, val_idx, val_rel, [len(y_pred)],
                                   T + len(extra) + 1, per_user=True)
    rec, nd, nd_b, rr = per_user[0].tolist()
    return rec, nd, nd_b, rr


def reciprocal_rank(y_true, y_pred):
    return _one_list({k: 1.0 for k in y_true}, y_pred)[3]


def ndcg(y_true, y_pred):
    return _one_list(y_true, y_pred)[1]


def binary_ndcg(y_true, y_pred):
    return _one_list({k: 1.0 for k in y_true}, y_pred)[2]


def binary_ndcg_v2(y_true, y_pred):
    return _one_list({k: 1.0 for k in y_true}, y_pred)[2]


def recall(y_true, y_pred):
    return _one_list({k: 1.0 for k in y_true}, y_pred)[0]


def average_precision(y_true, y_pred):
    """hwer/utils.py:81-98 (imported by hwer/validation.py:26, called nowhere in the reference): a repeated
    prediction only counts the first time."""
    y_pred = [p[0] if isinstance(p, (tuple, list, np.ndarray)) else p for p in y_pred]
    remaining = set(np.array(y_true).reshape(-1).tolist())
    len_y_true = max(1, len(y_true))
    hit = []
    for p in y_pred:
        hit.append(1.0 if p in remaining else 0.0)
        remaining.discard(p)
    if not hit:
        return 0.0
    h = torch.tensor(hit, dtype=torch.float64, device="cuda")
    ranks = torch.arange(1, len(hit) + 1, dtype=torch.float64, device="cuda")
    return float((h.cumsum(0) / ranks * h).sum().item()) / len_y_true

"""Host-side mirror of the hot-path helpers of hwer/utils.py, backed by the CUDA kernels.

Same names, argument meaning and return shapes as the reference so call sites read the same:
  unit_length(a, axis)                   hwer/utils.py:43-44 (+ repeat_args_wrapper :286-293,312)
  unit_length_violations(a, axis, eps)   hwer/utils.py:51-57
  NodeNotFoundException                  hwer/utils.py:326
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

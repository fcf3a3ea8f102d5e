"""Index-based point-cloud operators (reference: vgtk/vgtk/pc/sample.py:46-77)."""
import numpy as np
import torch

import vgtk.cuda.grouping as cuda_nn
import vgtk.utils as utils


def uniform_resample_index_np(pc, n_sample, batch=False):
    if batch:
        raise NotImplementedError('resample in batch is not implemented')
    n_point = pc.shape[0]
    if n_point >= n_sample:
        return np.random.choice(n_point, n_sample, replace=False)
    extra = np.random.choice(n_point, n_sample - n_point, replace=True)
    return np.concatenate((np.arange(n_point), extra), axis=0)


def uniform_resample_np(pc, n_sample, label=None, batch=False):
    idx = uniform_resample_index_np(pc, n_sample, batch)
    return (idx, pc[idx]) if label is None else (idx, pc[idx], label[idx])


def group_nd(pc, idx):
    """[b,c,n] x [b,m1(,m2,..)] -> [b,c,m1(,m2,..)]  (no autograd, like the reference)."""
    b = idx.shape[0]
    out = utils.batch_gather(pc, idx.reshape(b, -1).contiguous(), dim=2)
    return out.view(b, -1, *idx.shape[1:])


def ball_query_index(query_points, support_points, radius, n_sample):
    """[b,3,m] x [b,3,n] -> int32 [b,m,n_sample]"""
    return cuda_nn.ball_query(query_points, support_points, radius, n_sample)


_IDENTITY_INDEX = {}


def identity_index(nb, n, device, dtype):
    """[nb, n] with row = 0..n-1 (the "first n points" sample of lazy sampling / the sample index of a stride-1 layer).
    It depends on the shapes only, so it is built once per (shape, device, dtype) instead of by two or three small launches per
    layer and step.  The tensor is shared between callers: read-only."""
    key = (nb, n, str(device), dtype)
    t = _IDENTITY_INDEX.get(key)
    if t is None:
        t = torch.arange(n, device=device).view(1, -1).expand(nb, -1).to(dtype).contiguous()
        if not (t.is_cuda and torch.cuda.is_current_stream_capturing()):    # memory of a graph's private pool is not kept
            if len(_IDENTITY_INDEX) >= 64:
                _IDENTITY_INDEX.pop(next(iter(_IDENTITY_INDEX)))
            _IDENTITY_INDEX[key] = t
    return t


def furthest_sample_index(pc, n_sample, lazy_sample):
    """[b,3,n] -> int32 [b,n_sample]; `lazy_sample` (or n == n_sample) takes the first n_sample points."""
    if pc.shape[2] == n_sample or lazy_sample:
        return identity_index(pc.shape[0], n_sample, pc.device, torch.int32)
    return cuda_nn.furthest_point_sampling(pc, n_sample)


def furthest_sample(pc, n_sample, lazy_sample=True):
    idx = furthest_sample_index(pc, n_sample, lazy_sample)
    return idx, group_nd(pc, idx)

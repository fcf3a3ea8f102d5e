"""vgtk.utils (reference: vgtk/vgtk/utils.py)."""
import numpy as np
import torch

import vgtk.cuda.gathering as cuda_gather


def promote_input(x, n_dim, device=None):
    x = x.view((1,) * (n_dim - x.ndim) + tuple(x.shape))
    return x.to(device) if device is not None else x


def promote_input_np(x, n_dim):
    return x.reshape((1,) * (n_dim - x.ndim) + tuple(x.shape))


def batch_gather(x, idx, dim=1):
    """[b,c,n] x [b,m] -> [b,c,m]; like the reference (utils.py:25-27) this is NOT an autograd op."""
    return cuda_gather.gather_points_forward(x, idx.int())


def batch_zip(x, y, idx):
    raise NotImplementedError('batch zip cuda not implemented')


class LearningRateScheduler():
    """Step / exponential decay with optional warm-up (reference: vgtk/vgtk/utils.py:33-112)."""

    def __init__(self, optimizer, init_lr, lr_type, decay_step=None, decay_rate=None, **kwargs):
        self.optimizer, self.init_lr, self.lr, self.lr_type = optimizer, init_lr, init_lr, lr_type
        self.decay_step, self.decay_rate, self.counter = decay_step, decay_rate, 0
        self.kwargs = kwargs

    def _set(self, lr):
        self.lr = lr
        for g in self.optimizer.param_groups:
            g['lr'] = lr

    def step(self):
        self.counter += 1
        if self.lr_type in ('constant', None) or not self.decay_step:
            return
        if self.lr_type.startswith('exp') and self.counter % self.decay_step == 0:
            self._set(self.lr * (self.decay_rate or 1.0))
        elif self.lr_type.startswith('linear') and self.counter % self.decay_step == 0:
            self._set(max(self.lr - (self.decay_rate or 0.0), 0.0))

    def get_lr(self):
        return self.lr

import os
import random

import numpy as np
import torch

from .logger import Logger
from .summary import Summary
from .timer import Timer


class Trainer():
    """Skeleton of the training driver the reference's SPConvNets.trainer_* classes derive from
    (reference: vgtk/vgtk/app/trainer.py:17-224): seeding, run directories, logger, Adam + LR
    schedule, checkpoint save.  Subclasses supply _setup_datasets/_setup_model/_setup_metric/_optimize."""

    def __init__(self, opt):
        self.opt = opt
        self._set_seed(getattr(opt, 'seed', 0))
        root = os.path.join(getattr(opt, 'model_dir', './trained_models'), getattr(opt, 'experiment_id', 'exp'))
        self.root_dir = root
        os.makedirs(os.path.join(root, 'ckpt'), exist_ok=True)
        self.logger = Logger(os.path.join(root, 'log.txt'))
        self.summary = Summary()
        self.timer = Timer()
        self.iter_counter = 0
        self.epoch_counter = 0
        self._setup_datasets()
        self._setup_model()
        self._setup_optim()
        self._setup_metric()

    def _set_seed(self, seed):
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)

    def _setup_datasets(self):
        raise NotImplementedError

    def _setup_model(self):
        raise NotImplementedError

    def _setup_metric(self):
        pass

    def _setup_optim(self):
        lr = getattr(getattr(self.opt, 'train_lr', None), 'init_lr', 1e-3)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)

    def _optimize(self, data):
        raise NotImplementedError

    def train_iter(self):
        for _ in range(getattr(self.opt, 'num_iterations', 0)):
            self.timer.set_point()
            self.step()
            self.iter_counter += 1

    def step(self):
        raise NotImplementedError

    def _save_network(self, step, label=None, path=None):
        label = self.opt.experiment_id if label is None else label
        path = os.path.join(self.root_dir, 'ckpt', f'{label}_net_{step}.pth') if path is None else path
        state = self.model.module.state_dict() if hasattr(self.model, 'module') else self.model.state_dict()
        torch.save(state, path)
        return path

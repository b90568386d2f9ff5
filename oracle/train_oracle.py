"""CPU oracle of the TRAINING step of the same network -- TEST INFRASTRUCTURE ONLY.

Restates PointNet/train_semseg.py:164-179 (``classifier.train()``; forward; ``get_loss`` = weighted NLL,
models/pointnet2_sem_seg.py:43-49; ``loss.backward()``; ``optimizer.step()``) with the optimiser of :125-132
(Adam, betas (0.9, 0.999), eps 1e-8, L2 ``weight_decay`` added to the gradient) over the functional model of
oracle/pointnet2_oracle.py switched to train mode: BatchNorm uses batch statistics and updates its running
statistics in place, the head's Dropout(0.5) is active.

Parity status: PINNED against tests/golden/train_ssg.npz, which oracle/make_golden_train.py produced by running
the unmodified reference model + ``get_loss`` + ``torch.optim.Adam`` for two steps (tests/test_oracle_golden.py).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import pointnet2_oracle as PO


def param_keys(sd):
    """Trainable tensors of a checkpoint dict, in ``nn.Module.parameters()`` order of the reference model
    (= state_dict order without the BatchNorm buffers)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


class Trainer:
    """Holds the checkpoint dict (updated in place), the Adam state and the schedule of train_semseg.py."""

    def __init__(self, state_dict, arch="ssg", lr=1e-3, weight_decay=1e-4, bn_momentum=0.1):
        self.sd = OrderedDict((k, v.clone()) for k, v in state_dict.items())
        self.arch = arch
        self.keys = param_keys(self.sd)
        for k in self.keys:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.keys], lr=lr, betas=(0.9, 0.999), eps=1e-8,
                                    weight_decay=weight_decay)
        self.bn_momentum = bn_momentum

    def set_lr(self, lr):
        for g in self.opt.param_groups:
            g["lr"] = lr

    def loss_and_grads(self, x, target, weight=None, dropout_mask=None):
        """One train-mode forward + backward.  x [B,9,N]; target int64 [B,N]; weight [13] or None.
        Returns (loss, logp); gradients are left in ``.grad`` of the parameters."""
        PO.TRAIN = {"momentum": self.bn_momentum, "dropout_mask": dropout_mask}
        try:
            self.opt.zero_grad()
            logp, _ = PO.model_forward(self.sd, x, self.arch)
            pred = logp.contiguous().view(-1, logp.size(2))                     # train_semseg.py:170
            loss = F.nll_loss(pred, target.view(-1), weight=weight)             # pointnet2_sem_seg.py:47
            loss.backward()
        finally:
            PO.TRAIN = None
        return loss.detach(), logp.detach()

    def step(self, x, target, weight=None, dropout_mask=None):
        loss, logp = self.loss_and_grads(x, target, weight, dropout_mask)
        self.opt.step()
        return loss, logp

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.sd.items())

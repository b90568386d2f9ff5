"""Non-targeted colour attacks -- drop-in for PointNet/attacks/torchattacks/attacks/nontarget.py.

NB_attack (:10-42): L-inf sign ascent on sum-CE / N over channels 3:6; the returned tensor is the
un-projected last step (Q1).  The whole loop is one C call (psg_nb_attack): the geometry of all
``iters`` forwards is computed in one batched pass first, because for a colour attack it depends
only on xyz and on the FPS start draws (made here on the CPU generator in the reference's order).
"""
from __future__ import annotations

import torch

from ..attack import Attack

_MAX_PROBLEMS = 1024     # FPS problems (= forwards x blocks) whose geometry is resident at once


class NB_attack(Attack):
    def __init__(self, model, eps=0.3, alpha=2 / 255, iters=40):
        super().__init__("NB_attack", model)
        self.eps = eps
        self.alpha = alpha
        self.iters = iters

    def forward(self, images, labels):
        return _nb_loop(self, images, labels, target=-1, mask=None)


def _nb_loop(atk, images, labels, target, mask):
    eng = atk._engine(images)
    B, C, N = images.shape
    adv = images.detach().clone(memory_format=torch.contiguous_format)
    ori = images.detach()[:, 3:6].contiguous()
    lab = atk._labels_i32(labels, images.device) if target < 0 else None
    msk = atk._mask_u8(mask, B, N, images.device) if mask is not None else None
    # nontarget.py:34: sum-CE / N;  target.py:38: mean CE over the B*N points
    scale = 1.0 / N if target < 0 else 1.0 / (B * N)
    chunk = max(1, min(atk.iters, _MAX_PROBLEMS // B))
    eng.bind(B, N, chunk)
    eng.set_input(images.detach())
    done = 0
    while done < atk.iters:
        t = min(chunk, atk.iters - done)
        eng.geometry(eng.draw_starts(t))
        eng.nb_attack(adv, ori, msk, lab, target, t, 0, atk.alpha, atk.eps, scale)
        done += t
    atk.model._generation += 1
    return adv


class NU_attack(Attack):
    def __init__(self, model, c=1e-4, kappa=0, steps=1000, lr=0.01):
        super().__init__("NU_attack", model)
        self.c = c
        self.kappa = kappa
        self.steps = steps
        self.lr = lr

    def forward(self, images, labels):
        from pointsecguard_b200 import nu
        return nu.nu_attack(self, images, labels, mask=None, target=None, neighbour=10)

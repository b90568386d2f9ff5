#!/usr/bin/env python
"""Diagnostic: gradients w.r.t. the level features in the training backward, GPU engine vs the CPU training oracle."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pointnet2_oracle as PO, train_oracle as TO          # noqa: E402
from pointsecguard_b200 import synthetic as syn                       # noqa: E402
from pointsecguard_b200.models.pointnet2_sem_seg import get_model     # noqa: E402
from pointsecguard_b200.train import Trainer                          # noqa: E402


def main():
    torch.set_num_threads(8)
    sd = syn.make_state_dict("ssg", init="he")
    x, y = syn.make_painted_blocks(2, 1024, 50)
    torch.manual_seed(11)
    starts = [torch.randint(0, n, (2,), dtype=torch.long) for n in (1024, 1024, 256, 64)]
    keep = torch.empty(2, 128, 1024).bernoulli_(0.5)
    # ---- oracle with hooks on the level features
    tr = TO.Trainer(sd, "ssg")
    grads = {}
    orig_sa = PO.set_abstraction
    lvl = [0]

    def sa_hook(*a, **k):
        nx, nf = orig_sa(*a, **k)
        lvl[0] += 1
        i = lvl[0]
        nf.register_hook(lambda g, i=i: grads.__setitem__(i, g.detach().clone()))
        return nx, nf
    PO.set_abstraction = sa_hook
    it = iter(starts)
    orig_fps = PO.farthest_point_sample
    PO.farthest_point_sample = lambda xyz, n, start=None: orig_fps(xyz, n, next(it))
    loss, logp = tr.loss_and_grads(x, y, None, dropout_mask=keep)
    PO.set_abstraction, PO.farthest_point_sample = orig_sa, orig_fps
    # ---- GPU
    m = get_model(13); m.load_state_dict(sd); m = m.cuda()
    t = Trainer(m)
    t.eng.debug = {}
    l2, lp2 = t.loss_and_grads(x.cuda(), y.cuda(), None, dropout_mask=keep, starts=starts)
    print("loss", float(loss), float(l2), "logp maxdiff", (lp2.cpu() - logp).abs().max().item())
    for l in (4, 3, 2, 1):
        mine = t.eng.debug.get(l)
        ref = grads[l]                      # [B, C, S]
        if mine is None:
            continue
        d = (mine.cpu() - ref).abs().max().item()
        print(f"level {l}: d feats max |mine - ref| {d:.3e}, ref max {ref.abs().max().item():.3e}, shapes {tuple(mine.shape)} {tuple(ref.shape)}")
    byname = dict(m.named_parameters())
    for k in ("sa4.mlp_bns.2.bias", "sa4.mlp_bns.2.weight", "sa4.mlp_convs.2.weight", "sa4.mlp_bns.1.bias", "fp4.mlp_bns.0.bias"):
        a, b = t.grad_of(byname[k]).cpu(), tr.sd[k].grad
        print(k, "max|diff|", (a - b).abs().max().item(), "ref max", b.abs().max().item())


if __name__ == "__main__":
    main()

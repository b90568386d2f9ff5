#!/usr/bin/env python
"""Small workload compute-sanitizer runs over (tools/sanitize.sh): every kernel family of the attack path once or twice --
SSG and MSG, fp32 and tcgen05 modes, NB / tar-NB / NU (colour) / NU (coordinates + colour, geometric gradient), metrics."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from pointsecguard_b200 import metrics as MT, synthetic as syn, torchattacks      # noqa: E402
from pointsecguard_b200.engine import MLP_FP32, MLP_TF32                           # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for arch in ("ssg", "msg"):
        if arch == "ssg":
            from pointsecguard_b200.models.pointnet2_sem_seg import get_model
        else:
            from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
        m = get_model(13)
        m.load_state_dict(syn.make_state_dict(arch, init="he"))
        m = m.to(dev).eval()
        x, labels = syn.make_painted_blocks(2, 1024, 0)
        xd, lab = x.to(dev), labels.numpy().astype(np.float64)
        mask = labels == 11
        for mode in (MLP_FP32, MLP_TF32):
            m.set_mlp_mode(mode)
            torch.manual_seed(0)
            adv = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=2)(xd, lab)
            adv = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=2, target=7, mask=mask)(xd, lab)
            if arch == "ssg":
                torchattacks.NU_attack(m, c=0.1, kappa=0, steps=2, lr=0.01)(xd, lab)
                torchattacks.NU_attack(m, c=0.1, kappa=0, steps=2, lr=0.01, field=(0, 6))(xd, lab)
                torchattacks.tar_NU_attack(m, c=0.1, kappa=0, steps=2, lr=0.01, target=7, mask=mask[0].numpy())(xd[:1], lab[:1])
            logp, _ = m(adv)
            MT.attack_counters(logp, labels.to(dev), mask.to(dev), 7)
            xg = xd.clone().requires_grad_(True)
            m(xg)[0].sum().backward()
        torch.cuda.synchronize()
        print(arch, "ok", flush=True)


if __name__ == "__main__":
    main()

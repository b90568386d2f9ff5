#!/usr/bin/env python
"""Diagnostic: colour-gradient agreement between the GPU path (fp32 and TF32 modes) and the CPU oracle on the trained
checkpoint, at the clean input and at the golden trajectory's state before its last iteration (config 1)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pointnet2_oracle as PO                                    # noqa: E402
from pointsecguard_b200 import synthetic as syn                              # noqa: E402
from pointsecguard_b200.engine import MLP_FP32, MLP_TF32                     # noqa: E402
from pointsecguard_b200.models.pointnet2_sem_seg import get_model            # noqa: E402


def main():
    sd = syn.load_checkpoint("ssg")
    g = np.load(os.path.join(REPO, "tests", "golden", "atsize_config1.npz"))
    x, labels = syn.make_painted_blocks(4, 4096, 0)
    code = torch.from_numpy(g["prev"].astype(np.int16))
    col = x[:, 3:6] + code.float() * 0.05
    col = torch.where(code == -128, torch.zeros_like(col), torch.where(code == 127, torch.ones_like(col), col))
    x2 = x.clone()
    x2[:, 3:6] = col
    m = get_model(13); m.load_state_dict(sd); m = m.cuda().eval()
    om = PO.OracleModel(sd, "ssg")
    for name, inp in (("clean", x), ("step9", x2)):
        xo = inp.clone().requires_grad_(True)
        torch.manual_seed(3)
        lo, _ = om(xo)
        co = torch.nn.functional.cross_entropy(lo.reshape(-1, 13), labels.view(-1), reduction="sum") / 4096
        go, = torch.autograd.grad(co, xo)
        go = go[:, 3:6].numpy()
        for mode, mname in ((MLP_FP32, "fp32"), (MLP_TF32, "tf32")):
            m.set_mlp_mode(mode)
            xg = inp.cuda().requires_grad_(True)
            torch.manual_seed(3)
            lg, _ = m(xg)
            cg = torch.nn.functional.cross_entropy(lg.reshape(-1, 13), labels.cuda().view(-1), reduction="sum") / 4096
            gg, = torch.autograd.grad(cg, xg)
            gg = gg[:, 3:6].cpu().numpy()
            dl = (lg.detach().cpu() - lo.detach()).abs().max().item()
            rel = np.linalg.norm(gg - go) / np.linalg.norm(go)
            z_o, z_g = go == 0, gg == 0
            nz = ~z_o & ~z_g
            flips = (np.sign(gg[nz]) != np.sign(go[nz])).mean()
            amax = np.abs(go).max()
            small = np.abs(go[nz]) < 1e-6 * amax
            print(f"{name} {mname}: max|dlogp| {dl:.2e} grad rel {rel:.2e}; ref zeros {z_o.mean():.4f} mine zeros {z_g.mean():.4f} "
                  f"zero-mismatch {(z_o != z_g).mean():.5f}; sign flips among nonzero {flips:.5f}; |g|<1e-6 max: {small.mean():.4f}; "
                  f"quantiles |g|/max {np.quantile(np.abs(go[nz]) / amax, [0.01, 0.05, 0.25, 0.5])}")
            # the attack entry point on the same input and draws: one NB step = x + alpha * sign(g)
            from pointsecguard_b200 import torchattacks
            torch.manual_seed(3)
            adv1 = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=1)(inp.cuda(), labels.numpy().astype(np.float64))
            sg_attack = np.rint(((adv1[:, 3:6] - inp.cuda()[:, 3:6]) / 0.05).cpu().numpy())
            print(f"    attack step vs oracle sign: identical {(sg_attack == np.sign(go)).mean():.5f}; attack step vs autograd sign of the same mode: "
                  f"{(sg_attack == np.sign(gg)).mean():.5f}; moved where oracle gradient is zero: {(sg_attack[z_o] != 0).mean():.5f}")
            bad = nz.copy(); bad[nz] = np.sign(gg[nz]) != np.sign(go[nz])
            if bad.any():
                print("    flipped elements: median |g_ref|/max", np.median(np.abs(go[bad])) / amax, " median |g_mine - g_ref|/max", np.median(np.abs(gg[bad] - go[bad])) / amax)


if __name__ == "__main__":
    main()

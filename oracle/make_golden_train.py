"""Generate tests/golden/train_ssg.npz by executing the UNMODIFIED reference training step on the CPU.

Run in the build container only (needs /root/reference):      python -m oracle.make_golden_train

Two steps of PointNet/train_semseg.py:164-179 -- ``classifier.train()``, forward, ``get_loss`` (weighted NLL,
models/pointnet2_sem_seg.py:43-49), ``loss.backward()``, ``optimizer.step()`` with the Adam of :125-132 -- on
B=2 x 1024-point painted blocks, ``torch.manual_seed(11)`` before the loop (FPS starts and dropout masks come
from the global CPU generator in the reference's order).  Stored: the loss and log-probabilities of each step, the
gradients of a few tensors after step 1, per-tensor float64 sums of every gradient and of every checkpoint tensor
after step 2, and a few tensors in full.  Inputs and the initial checkpoint regenerate from seeds.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/PointNet"
sys.path[:0] = [REPO, REF, os.path.join(REF, "models")]
warnings.filterwarnings("ignore")

import pointnet2_sem_seg as RS                           # noqa: E402  (reference)
import pointnet2_sem_seg_msg as RM                       # noqa: E402  (reference)
from pointsecguard_b200 import synthetic as syn          # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
FULL = ["sa1.mlp_convs.0.weight", "sa1.mlp_bns.0.weight", "sa2.mlp_convs.2.bias", "sa4.mlp_bns.2.bias",
        "fp4.mlp_bns.0.weight", "fp1.mlp_convs.2.weight", "conv1.weight", "bn1.bias", "conv2.weight", "conv2.bias",
        "sa1.mlp_bns.0.running_mean", "sa1.mlp_bns.0.running_var", "bn1.running_var", "fp2.mlp_bns.1.running_mean"]
FULL_MSG = ["sa1.conv_blocks.0.0.weight", "sa1.bn_blocks.1.2.weight", "sa3.conv_blocks.1.1.weight", "fp4.mlp_bns.0.weight",
            "conv2.weight", "sa1.bn_blocks.0.0.running_mean", "sa2.bn_blocks.1.1.running_var", "bn1.running_var"]

CLASS_WEIGHTS = [1.0, 1.2, 0.8, 1.5, 1.0, 0.7, 1.3, 1.0, 0.9, 1.1, 1.4, 0.6, 1.0]


def case(arch, B=2, N=1024, steps=2):
    mod = RS if arch == "ssg" else RM
    m = mod.get_model(13)
    m.load_state_dict(syn.make_state_dict(arch, init="he"))
    crit = mod.get_loss()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    w = torch.tensor(CLASS_WEIGHTS)
    out = {}
    names = [k for k, _ in m.named_parameters()]
    torch.manual_seed(11)
    for s in range(steps):
        x, y = syn.make_painted_blocks(B, N, 50 + s)
        opt.zero_grad()
        m = m.train()
        pred, feat = m(x)
        loss = crit(pred.contiguous().view(-1, 13), y.view(-1), feat, w)
        loss.backward()
        out[f"loss{s}"] = np.float64(loss.item())
        out[f"logp{s}"] = pred.detach().numpy()
        out[f"gradsum{s}"] = np.array([p.grad.double().sum().item() for p in m.parameters()])
        out[f"gradabs{s}"] = np.array([p.grad.double().abs().sum().item() for p in m.parameters()])
        if s == 0:
            for k in (FULL if arch == "ssg" else FULL_MSG):
                if k in names:
                    out["grad0/" + k] = dict(m.named_parameters())[k].grad.numpy().copy()
        opt.step()
    sd = m.state_dict()
    out["pnames"] = np.array(names)          # nn.Module.parameters() order (gradsum / gradabs follow it)
    out["keys"] = np.array(list(sd.keys()))
    out["sum"] = np.array([v.double().sum().item() for v in sd.values()])
    out["abs"] = np.array([v.double().abs().sum().item() for v in sd.values()])
    for k in (FULL if arch == "ssg" else FULL_MSG):
        out["final/" + k] = sd[k].numpy().copy()
    out["rng_after"] = torch.get_rng_state().numpy()[:64].copy()
    return out


def main():
    torch.set_num_threads(1)          # fixed reduction order of the reference run
    for arch in ("ssg", "msg"):
        o = case(arch)
        path = os.path.join(OUT, f"train_{arch}.npz")
        np.savez_compressed(path, **o)
        print(arch, "loss", o["loss0"], o["loss1"], "->", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()

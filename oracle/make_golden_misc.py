"""tests/golden/api_misc.npz: outputs of the reference's ``sample_and_group_all`` (pointnet_util.py:146-163) and
``get_loss`` (pointnet2_sem_seg.py:43-49) on seeded inputs, made by executing the unmodified reference.
Build container only:      python -m oracle.make_golden_misc"""
import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/PointNet"
sys.path[:0] = [REPO, REF, os.path.join(REF, "models")]
warnings.filterwarnings("ignore")

import models.pointnet_util as R              # noqa: E402
import pointnet2_sem_seg as RS                # noqa: E402


def inputs():
    g = torch.Generator("cpu").manual_seed(31)
    xyz = torch.rand(2, 64, 3, generator=g)
    pts = torch.rand(2, 64, 5, generator=g)
    logits = torch.randn(2 * 64, 13, generator=g)
    target = torch.randint(0, 13, (2 * 64,), generator=g)
    weight = torch.rand(13, generator=g) + 0.5
    return xyz, pts, logits, target, weight


def main():
    xyz, pts, logits, target, weight = inputs()
    nx, npts = R.sample_and_group_all(xyz, pts)
    nx2, npts2 = R.sample_and_group_all(xyz, None)
    pred = torch.log_softmax(logits, 1).requires_grad_(True)
    loss = RS.get_loss()(pred, target, None, weight)
    loss.backward()
    out = {"sga_new_xyz": nx.numpy(), "sga_new_points": npts.numpy(), "sga_new_points_noattr": npts2.numpy(),
           "loss": np.float64(loss.item()), "dpred": pred.grad.numpy()}
    path = os.path.join(REPO, "tests", "golden", "api_misc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()

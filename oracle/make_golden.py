"""Generate tests/golden/*.npz by importing and executing the UNMODIFIED reference on the CPU.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python -m oracle.make_golden            # from the repo root

Inputs and checkpoints are regenerated from seeds by pointsecguard_b200.synthetic, so the goldens
hold outputs only.  The script also reports how the oracle restatement compares with the
reference on the same inputs (the pinned check itself lives in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import os
import sys
import warnings
import zlib

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/PointNet"
sys.path[:0] = [REPO, REF, os.path.join(REF, "models"), os.path.join(REF, "attacks")]
warnings.filterwarnings("ignore")

import models.pointnet_util as R       # noqa: E402  (reference; the module object the models use)
import pointnet2_sem_seg as RS                 # noqa: E402
import pointnet2_sem_seg_msg as RM             # noqa: E402
import torchattacks as RA                      # noqa: E402

from pointsecguard_b200 import synthetic as syn   # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
KINDS = ["uniform", "grid", "clustered", "duplicates", "surface"]


def crc(t):
    return np.uint32(zlib.crc32(np.ascontiguousarray(t.detach().numpy()).tobytes()))


def geom_case(kind, B, N, S, radii, nsamples, seed):
    x = syn.make_blocks(B, N, seed, kind)
    xyz = x[:, :3].permute(0, 2, 1).contiguous()
    torch.manual_seed(seed + 100)
    start = torch.randint(0, N, (B,), dtype=torch.long)
    torch.manual_seed(seed + 100)
    fps = R.farthest_point_sample(xyz, S)
    assert torch.equal(fps[:, 0], start)
    new_xyz = R.index_points(xyz, fps)
    out = {"start": start.numpy().astype(np.int32), "fps": fps.numpy().astype(np.int32)}
    for r, k in zip(radii, nsamples):
        out[f"ball_r{r}_k{k}"] = R.query_ball_point(r, k, xyz, new_xyz).numpy().astype(np.int32)
    d = R.square_distance(xyz, new_xyz)
    out["sqd_crc"] = crc(d)
    out["sqd_rows"] = d[:, :4].numpy()
    ds, idx = d.sort(dim=-1)
    ds, idx = ds[:, :, :3], idx[:, :, :3]
    rec = 1.0 / (ds + 1e-8)
    w = rec / torch.sum(rec, dim=2, keepdim=True)
    out["nn_idx"] = idx.numpy().astype(np.int32)
    out["nn_d2"] = ds.numpy()
    out["nn_w"] = w.numpy()
    return out


def load_ref(arch, seed=1234):
    m = (RS if arch == "ssg" else RM).get_model(13)
    m.load_state_dict(syn.make_state_dict(arch, seed))
    return m.eval()


def hooked_indices(model, x):
    """Record the FPS / ball-query / 3-NN indices of one reference forward by wrapping the
    reference's own functions (outputs only; nothing is altered)."""
    rec = []
    orig = (R.farthest_point_sample, R.query_ball_point)

    def fps(xyz, n):
        o = orig[0](xyz, n); rec.append(("fps", o.clone())); return o

    def ball(r, k, a, b):
        o = orig[1](r, k, a, b); rec.append(("ball", o.clone())); return o

    R.farthest_point_sample, R.query_ball_point = fps, ball
    try:
        out = model(x)
    finally:
        R.farthest_point_sample, R.query_ball_point = orig
    return out, rec


def model_case(arch, B, N):
    m = load_ref(arch)
    x = syn.make_blocks(B, N, 0, "uniform").clone().requires_grad_(True)
    torch.manual_seed(0)
    (logp, l4), rec = hooked_indices(m, x)
    labels = logp.detach().max(2)[1]
    # the NB cost of nontarget.py:34 with the clean prediction as label, shifted by one class so
    # that the gradient is not tiny everywhere
    y = (labels + 1) % 13
    cost = torch.nn.CrossEntropyLoss(reduction="sum")(logp.reshape(-1, 13), y.view(-1)) / logp.size(1)
    g, = torch.autograd.grad(cost, x)
    out = {"logp": logp.detach().numpy(), "l4": l4.detach().numpy(), "grad": g.numpy(),
           "y": y.numpy().astype(np.int8)}
    for i, (k, v) in enumerate(rec):
        out[f"idx{i:02d}_{k}"] = v.numpy().astype(np.int16)
    return out


def attack_cases():
    out = {}
    m = load_ref("ssg")
    # NB, B=2
    x = syn.make_blocks(2, 4096, 0, "uniform")
    torch.manual_seed(5)
    labels = m(x)[0].max(2)[1].numpy().astype(np.float64)
    torch.manual_seed(0)
    adv = RA.NB_attack(m, eps=0.1, alpha=0.05, iters=3)(x, labels)
    out["nb_labels"] = labels.astype(np.int8)
    out["nb_adv"] = adv.detach()[:, 3:6].numpy()
    assert torch.equal(adv.detach()[:, :3], x[:, :3]) and torch.equal(adv.detach()[:, 6:], x[:, 6:])
    # tar-NB, B=1, z-band labels, origin 11 -> target 7
    x1 = syn.make_blocks(1, 4096, 1, "uniform")
    zl = syn.zband_labels(x1)
    mask = (zl[0] == 11).numpy()
    torch.manual_seed(0)
    adv = RA.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=3, target=7, mask=mask)(x1, zl.numpy().astype(np.float64))
    out["tnb_adv"] = adv.detach()[:, 3:6].numpy()
    # NU, B=1
    torch.manual_seed(5)
    lab1 = m(x1)[0].max(2)[1].numpy().astype(np.float64)
    out["nu_labels"] = lab1.astype(np.int8)
    torch.manual_seed(0)
    adv = RA.NU_attack(m, c=0.1, kappa=0, steps=4, lr=0.01)(x1, lab1)
    out["nu_adv"] = adv.detach()[:, 3:6].numpy()
    # tar-NU, B=1, long enough to reach the stagnation test of target.py:127
    torch.manual_seed(0)
    adv = RA.tar_NU_attack(m, c=1, kappa=0, steps=22, lr=0.01, target=7, mask=mask)(x1, zl.numpy().astype(np.float64))
    out["tnu_adv"] = adv.detach().numpy()        # all nine channels: Q4 clamps xyz too
    # MSG NB, B=1
    mm = load_ref("msg")
    torch.manual_seed(5)
    labm = mm(x1)[0].max(2)[1].numpy().astype(np.float64)
    out["msg_nb_labels"] = labm.astype(np.int8)
    torch.manual_seed(0)
    adv = RA.NB_attack(mm, eps=0.1, alpha=0.05, iters=2)(x1, labm)
    out["msg_nb_adv"] = adv.detach()[:, 3:6].numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for kind in KINDS:
        np.savez_compressed(os.path.join(OUT, f"geom_{kind}.npz"),
                            **geom_case(kind, 2, 1024, 256, [0.2, 0.1], [32, 16], 3))
        print("geom", kind)
    np.savez_compressed(os.path.join(OUT, "geom_sa1.npz"),
                        **geom_case("uniform", 1, 4096, 1024, [0.1, 0.05], [32, 16], 4))
    for arch in ("ssg", "msg"):
        np.savez_compressed(os.path.join(OUT, f"model_{arch}.npz"), **model_case(arch, 2, 2048))
        print("model", arch)
    np.savez_compressed(os.path.join(OUT, "attack.npz"), **attack_cases())
    print("attacks")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

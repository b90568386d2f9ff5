"""CPU oracle: functional PyTorch restatement of the reference's PointNet++ sem-seg forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is pure Python over stock
PyTorch; this file restates the same computation as plain functions over a checkpoint dict, so it
can travel to the GPU box where /root/reference does not exist.  Every function cites the
reference lines it follows.  It runs on the CPU with autograd providing the input gradient, as the
reference does.

Parity status: PINNED.  tests/test_oracle_golden.py checks these functions against
tests/golden/*.npz, which oracle/make_golden.py produced by importing and executing the
unmodified reference in the build container.

geometry="torch" follows the reference op for op (matmul + sort); geometry="c" routes the four
index primitives through oracle/geom_oracle.c (same results, bit for bit on the golden vectors,
and much faster at large N).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import geom as cgeom

GEOMETRY = "c"   # default backend of the index primitives inside the model oracle


# --------------------------------------------------------------------------------------------
# primitives  (PointNet/models/pointnet_util.py:19-107)
# --------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """pointnet_util.py:19-40: -2 src.dst^T, then += |src|^2, then += |dst|^2 (in place)."""
    d = -2 * torch.matmul(src, dst.transpose(1, 2))
    d += (src ** 2).sum(-1).unsqueeze(2)
    d += (dst ** 2).sum(-1).unsqueeze(1)
    return d


def index_points(points, idx):
    """pointnet_util.py:43-60: points[b, idx[b, ...], :]."""
    B = points.shape[0]
    bsel = torch.arange(B, device=points.device).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[bsel, idx]


def farthest_point_sample(xyz, npoint, start=None):
    """pointnet_util.py:63-84.  ``start`` None draws torch.randint(0, N, (B,)) on the global CPU
    generator exactly like line 75."""
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)         # CPU generator, then moved (pointnet_util.py:75)
    if GEOMETRY == "c":
        return cgeom.fps(xyz, npoint, start)
    dev = xyz.device                                                # (device-agnostic: tools/stock_torch_baseline.py runs it on a GPU)
    out = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    mind = torch.full((B, N), 1e10, device=dev)
    far = start.clone().to(dev)
    rows = torch.arange(B, device=dev)
    for i in range(npoint):
        out[:, i] = far
        c = xyz[rows, far].unsqueeze(1)
        d = ((xyz - c) ** 2).sum(-1)
        mind = torch.where(d < mind, d, mind)
        far = mind.max(-1)[1]
    return out


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet_util.py:87-107."""
    if GEOMETRY == "c":
        return cgeom.ball_query(radius, nsample, xyz.detach(), new_xyz.detach())
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    d = square_distance(new_xyz, xyz)
    idx = torch.arange(N, device=xyz.device).expand(B, S, N).clone()
    idx[d > radius ** 2] = N
    idx = idx.sort(-1)[0][:, :, :nsample]
    first = idx[:, :, :1].expand(-1, -1, nsample)
    return torch.where(idx == N, first, idx)


def three_nn(xyz1, xyz2):
    """pointnet_util.py:301-303: three smallest expansion-form distances, (d2, idx)."""
    if GEOMETRY == "c":
        idx, d2, _ = cgeom.three_nn(xyz1.detach(), xyz2.detach())
        if xyz1.requires_grad or xyz2.requires_grad:
            # forward value from the C oracle, derivative of the same expansion formula through
            # autograd (the reference lets gradient reach xyz through the weights)
            nb = index_points(xyz2, idx)                                   # [B,N,3,3]
            dd = -2 * (xyz1.unsqueeze(2) * nb).sum(-1) + (xyz1 ** 2).sum(-1, keepdim=True) + (nb ** 2).sum(-1)
            d2 = d2 + (dd - dd.detach())
        return d2, idx
    d, idx = square_distance(xyz1, xyz2).sort(-1)
    return d[:, :, :3], idx[:, :, :3]


# --------------------------------------------------------------------------------------------
# layers  (pointnet_util.py:166-320).  Eval mode unless TRAIN is set (oracle/train_oracle.py): then BatchNorm
# normalises with the batch statistics and updates the running ones in place in ``sd`` (momentum TRAIN["momentum"]),
# as nn.BatchNorm{1,2}d.train() does, and the head's Dropout(0.5) is active.
# --------------------------------------------------------------------------------------------
TRAIN = None


def _conv_bn_relu(sd, cp, bp, x):
    w, b = sd[cp + ".weight"], sd[cp + ".bias"]
    y = F.conv2d(x, w, b) if w.dim() == 4 else F.conv1d(x, w, b)
    if TRAIN is not None:
        sd[bp + ".num_batches_tracked"] += 1
        y = F.batch_norm(y, sd[bp + ".running_mean"], sd[bp + ".running_var"],
                         sd[bp + ".weight"], sd[bp + ".bias"], True, TRAIN["momentum"], 1e-5)
    else:
        y = F.batch_norm(y, sd[bp + ".running_mean"], sd[bp + ".running_var"],
                         sd[bp + ".weight"], sd[bp + ".bias"], False, 0.0, 1e-5)
    return F.relu(y)


def set_abstraction(sd, name, arch, npoint, radii, nsamples, n_mlp, xyz, points, trace=None):
    """SSG: pointnet_util.py:181-207 (+ sample_and_group :110-143); MSG: :229-267.
    xyz [B,3,N], points [B,D,N] -> new_xyz [B,3,S], new_points [B,D',S]."""
    xyz_t = xyz.permute(0, 2, 1)
    pts_t = points.permute(0, 2, 1)
    B = xyz_t.shape[0]
    fps_idx = farthest_point_sample(xyz_t.detach(), npoint)
    new_xyz = index_points(xyz_t, fps_idx)
    if trace is not None:
        trace.append(("fps", fps_idx))
    outs = []
    for bi, (radius, K) in enumerate(zip(radii, nsamples)):
        gidx = query_ball_point(radius, K, xyz_t, new_xyz)
        if trace is not None:
            trace.append(("ball", gidx))
        gxyz = index_points(xyz_t, gidx) - new_xyz.view(B, npoint, 1, 3)
        gpts = index_points(pts_t, gidx)
        if arch == "ssg":
            g = torch.cat([gxyz, gpts], -1)          # :137  xyz first
        else:
            g = torch.cat([gpts, gxyz], -1)          # :253  features first
        g = g.permute(0, 3, 2, 1)                    # [B, C, K, S]
        for j in range(n_mlp[bi]):
            if arch == "ssg":
                g = _conv_bn_relu(sd, f"{name}.mlp_convs.{j}", f"{name}.mlp_bns.{j}", g)
            else:
                g = _conv_bn_relu(sd, f"{name}.conv_blocks.{bi}.{j}", f"{name}.bn_blocks.{bi}.{j}", g)
        outs.append(g.max(2)[0])
    return new_xyz.permute(0, 2, 1), torch.cat(outs, 1)


def feature_propagation(sd, name, n_mlp, xyz1, xyz2, points1, points2, trace=None):
    """pointnet_util.py:281-320 (the S == 1 branch is never taken on this path)."""
    x1 = xyz1.permute(0, 2, 1)
    x2 = xyz2.permute(0, 2, 1)
    p2 = points2.permute(0, 2, 1)
    B, N, _ = x1.shape
    d, idx = three_nn(x1, x2)
    if trace is not None:
        trace.append(("nn3", idx))
    recip = 1.0 / (d + 1e-8)
    w = recip / recip.sum(2, keepdim=True)
    interp = (index_points(p2, idx) * w.view(B, N, 3, 1)).sum(2)
    if points1 is not None:
        feat = torch.cat([points1.permute(0, 2, 1), interp], -1)
    else:
        feat = interp
    feat = feat.permute(0, 2, 1)
    for j in range(n_mlp):
        feat = _conv_bn_relu(sd, f"{name}.mlp_convs.{j}", f"{name}.mlp_bns.{j}", feat)
    return feat


# --------------------------------------------------------------------------------------------
# models  (pointnet2_sem_seg.py:22-40, pointnet2_sem_seg_msg.py:23-41)
# --------------------------------------------------------------------------------------------
_SA = {
    "ssg": [(1024, [0.1], [32]), (256, [0.2], [32]), (64, [0.4], [32]), (16, [0.8], [32])],
    "msg": [(1024, [0.05, 0.1], [16, 32]), (256, [0.1, 0.2], [16, 32]),
            (64, [0.2, 0.4], [16, 32]), (16, [0.4, 0.8], [16, 32])],
}
_FP_LAYERS = [2, 2, 2, 3]      # fp4, fp3, fp2, fp1


def model_forward(sd, x, arch="ssg", trace=None):
    """x [B,9,N] -> (log-probabilities [B,N,13], l4_points)."""
    xyz0 = x[:, :3, :]
    feats = [x]
    xyzs = [xyz0]
    for li, (npoint, radii, nsamples) in enumerate(_SA[arch]):
        nx, nf = set_abstraction(sd, f"sa{li+1}", arch, npoint, radii, nsamples,
                                 [3] * len(radii), xyzs[-1], feats[-1], trace)
        xyzs.append(nx)
        feats.append(nf)
    l4_points = feats[4]
    up = feats[4]
    up = feature_propagation(sd, "fp4", 2, xyzs[3], xyzs[4], feats[3], up, trace)
    up = feature_propagation(sd, "fp3", 2, xyzs[2], xyzs[3], feats[2], up, trace)
    up = feature_propagation(sd, "fp2", 2, xyzs[1], xyzs[2], feats[1], up, trace)
    up = feature_propagation(sd, "fp1", 3, xyzs[0], xyzs[1], None, up, trace)
    h = _conv_bn_relu(sd, "conv1", "bn1", up)          # dropout is the identity in eval mode
    if TRAIN is not None:
        if TRAIN.get("dropout_mask") is not None:      # injected keep-mask [B,128,N] (0/1), scaled like F.dropout
            h = h * TRAIN["dropout_mask"] * 2.0
        else:
            h = F.dropout(h, 0.5, True)                # pointnet2_sem_seg.py:18,35
    z = F.conv1d(h, sd["conv2.weight"], sd["conv2.bias"])
    logp = F.log_softmax(z, dim=1).permute(0, 2, 1)
    return logp, l4_points


class OracleModel:
    """Callable with the reference model's call shape: model(x) -> (logp, l4_points)."""

    def __init__(self, state_dict, arch="ssg"):
        self.sd = {k: v.clone() for k, v in state_dict.items()}
        self.arch = arch
        self.trace = None

    def __call__(self, x):
        return model_forward(self.sd, x, self.arch, self.trace)

"""Drop-in for PointNet/models/pointnet_util.py of the reference (same names, signatures, tensor
conventions and state_dict keys), backed by libpsg_b200.so.  CUDA tensors only.

Functions take [B, N, C] tensors, modules take / return channel-first [B, C, N] tensors, indices are
``torch.long`` -- exactly as in the reference (pointnet_util.py:19-320).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from pointsecguard_b200 import ops as _ops  # noqa: F401  (registers torch.ops.psg.*)


def pc_normalize(pc):
    """pointnet_util.py:11-17 (numpy helper, kept for API completeness)."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    return pc / m


def square_distance(src, dst):
    """pointnet_util.py:19-40.  src [B,N,C], dst [B,M,C] -> [B,N,M] in the reference's expansion
    form ((-2 src.dst) + |src|^2) + |dst|^2, bit-identical to the stock CPU path for C == 3."""
    if src.shape[-1] != 3:
        raise NotImplementedError("square_distance: only 3-D coordinates are on the hot path")
    return torch.ops.psg.square_distance(src, dst)


def index_points(points, idx):
    """pointnet_util.py:43-60.  points [B,N,C], idx [B,S] or [B,S,K] (long) -> [B,S,(K,)C]."""
    return _IndexPoints.apply(points, idx)


def farthest_point_sample(xyz, npoint):
    """pointnet_util.py:63-84.  The start index is drawn with torch.randint on the global *CPU*
    generator, exactly like line 75, so a seeded run reproduces the reference's indices."""
    B, N, _ = xyz.shape
    start = torch.randint(0, N, (B,), dtype=torch.long)
    return torch.ops.psg.fps(xyz.detach(), int(npoint), start)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet_util.py:87-107.  -> group_idx [B,S,nsample] (long)."""
    return torch.ops.psg.ball_query(float(radius), int(nsample), xyz.detach(), new_xyz.detach())


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """pointnet_util.py:110-143 (without its five torch.cuda.empty_cache() calls)."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        grouped_points = index_points(points, idx)
        new_points = torch.cat([grouped_xyz_norm, grouped_points], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet_util.py:146-163.  Never executed by the sem-seg networks (group_all=False in every
    layer); kept for signature compatibility as plain tensor reshapes."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device, dtype=xyz.dtype)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points


class _IndexPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.shape = points.shape
        return torch.ops.psg.index_points(points, idx)

    @staticmethod
    def backward(ctx, g):
        from pointsecguard_b200 import modules as _m
        idx, = ctx.saved_tensors
        return _m.index_points_backward(g, idx, ctx.shape), None


class PointNetSetAbstraction(nn.Module):
    """pointnet_util.py:166-207.  Parameter containers are the same nn.Conv2d / nn.BatchNorm2d
    modules as in the reference, so checkpoints load unchanged."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.group_all = group_all

    def forward(self, xyz, points):
        from pointsecguard_b200 import modules as _m
        return _m.set_abstraction_forward(self, xyz, points)


class PointNetSetAbstractionMsg(nn.Module):
    """pointnet_util.py:210-267."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint = npoint
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for i in range(len(mlp_list)):
            convs = nn.ModuleList()
            bns = nn.ModuleList()
            last_channel = in_channel + 3
            for out_channel in mlp_list[i]:
                convs.append(nn.Conv2d(last_channel, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        from pointsecguard_b200 import modules as _m
        return _m.set_abstraction_msg_forward(self, xyz, points)


class PointNetFeaturePropagation(nn.Module):
    """pointnet_util.py:270-320."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        from pointsecguard_b200 import modules as _m
        return _m.feature_propagation_forward(self, xyz1, xyz2, points1, points2)

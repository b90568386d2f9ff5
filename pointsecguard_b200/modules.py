"""Stand-alone forward / input-gradient backward of the three PointNet++ modules, composed from the
T-layout building blocks of the C ABI (include/psg_b200.h): group -> shared MLP -> neighbourhood
max (set abstraction, SSG and MSG) and 3-NN interpolate -> concat -> MLP (feature propagation).

Reference: PointNet/models/pointnet_util.py:166-207 (PointNetSetAbstraction.forward), :210-267
(PointNetSetAbstractionMsg.forward), :270-320 (PointNetFeaturePropagation.forward).

The sem-seg networks do not come through here -- ``get_model.forward`` runs the whole network
inside one ``psg_net`` (engine.py) -- but the module classes are part of the reference's API
surface and are usable on their own with the same tensor conventions (channel-first I/O, long
indices, eval-mode BatchNorm folded into the 1x1 convs).  Gradients flow to ``points`` (features);
the sampling / grouping indices are non-differentiable exactly as in the reference, and the
geometric gradient w.r.t. ``xyz`` (through ``grouped_xyz_norm`` and the interpolation weights) is
not produced here (the colour attacks never use it).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .engine import MLP_FP32, fold_conv_bn
from .tlayout import TTensor


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda_eval(module, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pointsecguard_b200 modules run on CUDA tensors only; there is no CPU fallback")
    if module.training:
        raise RuntimeError("pointsecguard_b200 implements the eval-mode (attack) path only; call module.eval()")


class _Chain:
    """psg_mlp handles of one conv+BN chain, rebuilt whenever a parameter changes."""

    def __init__(self, convs, bns, xyz_first: bool):
        self.handles, self.kpad, self.npad, self.cout = [], [], [], []
        for j, (conv, bn) in enumerate(zip(convs, bns)):
            w, b = fold_conv_bn(conv.weight, conv.bias,
                                {"weight": bn.weight, "bias": bn.bias, "running_mean": bn.running_mean,
                                 "running_var": bn.running_var}, bn.eps)
            if j == 0 and xyz_first:     # SSG groups [xyz_norm | feats] (:137); the library [feats | xyz_norm]
                w = w[:, list(range(3, w.shape[1])) + [0, 1, 2]]
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            h = L.psg_mlp_create(w.ctypes.data, b.ctypes.data, w.shape[1], w.shape[0])
            if not h:
                raise L.PsgError("psg_mlp_create failed")
            self.handles.append(h)
            self.kpad.append((w.shape[1] + 15) // 16 * 16)
            self.npad.append((w.shape[0] + 15) // 16 * 16)
            self.cout.append(w.shape[0])

    def __del__(self):
        destroy = getattr(L, "psg_mlp_destroy", None)
        for h in getattr(self, "handles", []):
            if destroy is not None:
                destroy(h)
        self.handles = []


def _chains(module, blocks, xyz_first):
    key = tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))
    cache = module.__dict__.get("_psg_chains")
    if cache is None or cache[0] != key:
        with torch.cuda.device(next(module.parameters()).device):
            cache = (key, [_Chain(c, b, xyz_first) for c, b in blocks])
        module.__dict__["_psg_chains"] = cache
    return cache[1]


def _mode(module):
    return getattr(module, "mlp_mode", MLP_FP32)


def _chain_forward(chain, a1, k1chunks, a2, k2chunks, rows, device, mode):
    ys = []
    for j, h in enumerate(chain.handles):
        out = TTensor(rows, chain.cout[j], device)
        L.psg_mlp_forward(h, a1.ptr, a1.wchunks, 0, k1chunks, a2.ptr if a2 is not None else None,
                          a2.wchunks if a2 is not None else 0, 0, k2chunks, rows, out.ptr, out.wchunks, 1, mode, _stream())
        ys.append(out)
        a1, k1chunks, a2, k2chunks = out, out.wchunks, None, 0
    return ys


def _chain_backward(chain, ys, dy, rows, device, mode):
    """dy: gradient w.r.t. the pre-activation of the last layer -> gradient w.r.t. the chain input."""
    cur = dy
    for j in range(len(chain.handles) - 1, -1, -1):
        dx = TTensor(rows, chain.kpad[j], device)
        mask = ys[j - 1] if j > 0 else None
        L.psg_mlp_backward(chain.handles[j], cur.ptr, cur.wchunks, rows, dx.ptr, dx.wchunks,
                           mask.ptr if mask is not None else None, mask.wchunks if mask is not None else 0, mode, _stream())
        cur = dx
    return cur


def _csr(keys_i32, P, M, R, device, pad_group=0):
    offs = torch.empty(P * (R + 1), dtype=torch.int32, device=device)
    perm = torch.empty(P * M, dtype=torch.int32, device=device)
    ws = torch.empty(max(L.psg_csr_workspace(P, M, R), 16), dtype=torch.uint8, device=device)
    L.psg_csr_build_by_source(keys_i32.data_ptr(), P, M, R, pad_group, offs.data_ptr(), perm.data_ptr(), ws.data_ptr(),
                              _stream())
    return offs, perm


# --------------------------------------------------------------------------------------------------
# set abstraction (SSG and MSG)
# --------------------------------------------------------------------------------------------------
class _SetAbstractionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, xyz_t, new_xyz, idxs, chains, mode):
        """points [B,D,N] (channel-first, any strides) or None; xyz_t [B,N,3]; new_xyz [B,S,3];
        idxs: int32 [B,S,K] per branch -> [B, sum(Cout), S]"""
        dev = xyz_t.device
        B, N, _ = xyz_t.shape
        S = new_xyz.shape[1]
        if points is not None:
            D = points.shape[1]
            feats = TTensor.from_channels_first(points.detach())
        else:
            D = 0
            feats = TTensor(B * N, 16, dev, zero=True)
        ctot = sum(c.cout[-1] for c in chains)
        out = TTensor(B * S, ctot, dev)
        saved = []
        col = 0
        for idx, chain in zip(idxs, chains):
            K = idx.shape[2]
            if chain.cout[-1] % 16:
                raise L.PsgError("set abstraction: branch output widths must be multiples of 16")
            rows = B * S * K
            G = TTensor(rows, D + 3, dev)
            L.psg_group_points(feats.ptr, feats.wchunks, D, xyz_t.data_ptr(), B, N, new_xyz.data_ptr(), idx.data_ptr(), B, S,
                               K, G.ptr, G.cpad, _stream())
            ys = _chain_forward(chain, G, G.wchunks, None, 0, rows, dev, mode)
            arg = torch.empty(B * S * ys[-1].cpad, dtype=torch.uint8, device=dev)
            L.psg_group_max(ys[-1].ptr, ys[-1].wchunks, B * S, K, ys[-1].cpad, out.ptr, out.wchunks, col // 4,
                            arg.data_ptr(), _stream())
            saved.append((idx, chain, ys, arg, col, K))
            col += chain.cout[-1]
        ctx.saved = (saved, out, B, N, S, D, mode)
        return out.to_channels_first(B, S, ctot)

    @staticmethod
    @L.on_device_of
    def backward(ctx, dout):
        saved, out, B, N, S, D, mode = ctx.saved
        if D == 0:
            return None, None, None, None, None, None
        dev = dout.device
        dT = TTensor.from_channels_first(dout)
        dfeats = TTensor(B * N, D, dev)
        for bi, (idx, chain, ys, arg, col, K) in enumerate(saved):
            rows = B * S * K
            cw = ys[-1].cpad
            dy = TTensor(rows, cw, dev)
            L.psg_group_max_backward(dT.ptr, dT.wchunks, col // 4, out.ptr, out.wchunks, col // 4, arg.data_ptr(), B * S, K,
                                     cw, dy.ptr, dy.wchunks, _stream())
            dG = _chain_backward(chain, ys, dy, rows, dev, mode)
            offs, perm = _csr(idx, B, S * K, N, dev, pad_group=K)
            L.psg_segment_sum(dG.ptr, dG.wchunks, 0, S * K, 1, None, offs.data_ptr(), perm.data_ptr(), S * K, N, B, D,
                              dfeats.ptr, dfeats.wchunks, 0, 1 if bi > 0 else 0, _stream())
        return dfeats.to_channels_first(B, N, D), None, None, None, None, None


def _sample(module, xyz):
    """FPS (start drawn on the CPU generator like pointnet_util.py:75) + the sampled coordinates."""
    from .models.pointnet_util import farthest_point_sample
    xyz_t = xyz.detach().permute(0, 2, 1).contiguous()
    fps_idx = farthest_point_sample(xyz_t, module.npoint)
    new_xyz = torch.ops.psg.index_points(xyz_t, fps_idx)
    return xyz_t, new_xyz


def _ball_i32(xyz_t, new_xyz, radii, ks):
    import ctypes as C
    B, N, _ = xyz_t.shape
    S = new_xyz.shape[1]
    outs = []
    for i in range(0, len(radii), 2):            # radii share a scan two at a time (MSG)
        r, k = list(radii[i:i + 2]), list(ks[i:i + 2])
        o = [torch.empty(B, S, kk, dtype=torch.int32, device=xyz_t.device) for kk in k]
        ra = (C.c_double * 2)(*(r + [0.0])[:2])
        ka = (C.c_int * 2)(*(k + [0])[:2])
        L.psg_ball_query(xyz_t.data_ptr(), B, B, N, new_xyz.data_ptr(), S, len(r), ra, ka, o[0].data_ptr(),
                         o[1].data_ptr() if len(o) > 1 else None, _stream())
        outs += o
    return outs


@L.on_device_of
def set_abstraction_forward(module, xyz, points):
    """PointNetSetAbstraction.forward, pointnet_util.py:181-207: xyz [B,3,N], points [B,D,N] ->
    (new_xyz [B,3,S], new_points [B,Cout,S])."""
    _need_cuda_eval(module, xyz, points)
    if module.group_all:
        raise NotImplementedError("group_all=True is not on the sem-seg attack path (SURVEY.md a6)")
    for k in (module.nsample,):
        if k not in (16, 32):
            raise L.PsgError("set abstraction: nsample must be 16 or 32")
    chains = _chains(module, [(module.mlp_convs, module.mlp_bns)], xyz_first=True)
    xyz_t, new_xyz = _sample(module, xyz)
    idxs = _ball_i32(xyz_t, new_xyz, [module.radius], [module.nsample])
    new_points = _SetAbstractionFn.apply(points, xyz_t, new_xyz, idxs, chains, _mode(module))
    return new_xyz.permute(0, 2, 1), new_points


@L.on_device_of
def set_abstraction_msg_forward(module, xyz, points):
    """PointNetSetAbstractionMsg.forward, pointnet_util.py:229-267 (features first, xyz last, :253)."""
    _need_cuda_eval(module, xyz, points)
    for k in module.nsample_list:
        if k not in (16, 32):
            raise L.PsgError("set abstraction: nsample must be 16 or 32")
    blocks = [(module.conv_blocks[i], module.bn_blocks[i]) for i in range(len(module.radius_list))]
    chains = _chains(module, blocks, xyz_first=False)
    xyz_t, new_xyz = _sample(module, xyz)
    idxs = _ball_i32(xyz_t, new_xyz, list(module.radius_list), list(module.nsample_list))
    new_points = _SetAbstractionFn.apply(points, xyz_t, new_xyz, idxs, chains, _mode(module))
    return new_xyz.permute(0, 2, 1), new_points


# --------------------------------------------------------------------------------------------------
# feature propagation
# --------------------------------------------------------------------------------------------------
class _FeaturePropagationFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points1, points2, nn_idx, nn_w, chain, mode):
        """points1 [B,D1,N] or None, points2 [B,D2,S]; nn_idx int32 / nn_w float32 [B,N,3]"""
        dev = points2.device
        B, D2, S = points2.shape
        N = nn_idx.shape[1]
        rows = B * N
        p2 = TTensor.from_channels_first(points2.detach())
        D1 = points1.shape[1] if points1 is not None else 0
        two_source = D1 > 0 and D1 % 16 == 0 and D2 % 16 == 0
        if D1 == 0 or two_source:
            interp = TTensor(rows, D2, dev)
            L.psg_interpolate(p2.ptr, p2.wchunks, S, nn_idx.data_ptr(), nn_w.data_ptr(), B, N, p2.cpad, interp.ptr,
                              interp.wchunks, 0, _stream())
            if D1 == 0:
                a1, k1, a2, k2 = interp, interp.wchunks, None, 0
            else:
                a1 = TTensor.from_channels_first(points1.detach())
                k1, a2, k2 = a1.wchunks, interp, interp.wchunks
        else:
            # widths that do not fall on 16-column boundaries: materialise the concatenation
            interp = TTensor(rows, D2, dev)
            L.psg_interpolate(p2.ptr, p2.wchunks, S, nn_idx.data_ptr(), nn_w.data_ptr(), B, N, p2.cpad, interp.ptr,
                              interp.wchunks, 0, _stream())
            cat = torch.cat([points1.detach(), interp.to_channels_first(B, N, D2)], dim=1)
            a1 = TTensor.from_channels_first(cat)
            k1, a2, k2 = a1.wchunks, None, 0
        ys = _chain_forward(chain, a1, k1, a2, k2, rows, dev, mode)
        ctx.saved = (chain, ys, nn_idx, nn_w, B, N, S, D1, D2, two_source, mode,
                     a1.cpad if D1 else 0)
        return ys[-1].to_channels_first(B, N, chain.cout[-1])

    @staticmethod
    @L.on_device_of
    def backward(ctx, dout):
        chain, ys, nn_idx, nn_w, B, N, S, D1, D2, two_source, mode, c1pad = ctx.saved
        dev = dout.device
        rows = B * N
        # gradient w.r.t. the pre-activation of the last layer: ReLU mask of its own output
        dy = TTensor.from_channels_first(dout * (ys[-1].to_channels_first(B, N, chain.cout[-1]) > 0))
        dcat = _chain_backward(chain, ys, dy, rows, dev, mode)
        d1 = None
        if D1:
            d1 = dcat.to_channels_first(B, N, D1, 0)
        # interpolated columns start after the (padded, in the two-source layout) skip columns
        c0 = c1pad if two_source else D1
        offs, perm = _csr(nn_idx, B, N * 3, S, dev)
        d2 = TTensor(B * S, D2, dev)
        if c0 % 4 == 0:
            L.psg_segment_sum(dcat.ptr, dcat.wchunks, c0 // 4, N, 3, nn_w.data_ptr(), offs.data_ptr(), perm.data_ptr(), N * 3,
                              S, B, D2, d2.ptr, d2.wchunks, 0, 0, _stream())
        else:
            sl = TTensor.from_channels_first(dcat.to_channels_first(B, N, D1 + D2, 0)[:, D1:])
            L.psg_segment_sum(sl.ptr, sl.wchunks, 0, N, 3, nn_w.data_ptr(), offs.data_ptr(), perm.data_ptr(), N * 3, S, B, D2,
                              d2.ptr, d2.wchunks, 0, 0, _stream())
        return d1, d2.to_channels_first(B, S, D2), None, None, None, None


@L.on_device_of
def feature_propagation_forward(module, xyz1, xyz2, points1, points2):
    """PointNetFeaturePropagation.forward, pointnet_util.py:281-320: xyz1 [B,3,N], xyz2 [B,3,S],
    points1 [B,D1,N] or None, points2 [B,D2,S] -> [B,Cout,N]."""
    _need_cuda_eval(module, xyz1, xyz2, points1, points2)
    chain = _chains(module, [(module.mlp_convs, module.mlp_bns)], xyz_first=False)[0]
    x1 = xyz1.detach().permute(0, 2, 1).contiguous()
    x2 = xyz2.detach().permute(0, 2, 1).contiguous()
    B, N, _ = x1.shape
    S = x2.shape[1]
    dev = x1.device
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=dev)
    w = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    if S == 1:
        # :298-299: a single coarse point is repeated for every fine point
        idx.zero_()
        w.zero_()
        w[..., 0] = 1.0
    elif S < 3:
        raise L.PsgError("feature propagation needs S == 1 or S >= 3 coarse points")
    else:
        L.psg_three_nn(x1.data_ptr(), B, B, N, x2.data_ptr(), S, idx.data_ptr(), w.data_ptr(), None, _stream())
    return _FeaturePropagationFn.apply(points1, points2, idx, w, chain, _mode(module))


# --------------------------------------------------------------------------------------------------
# index_points backward (autograd of pointnet_util.py:43-60), deterministic
# --------------------------------------------------------------------------------------------------
@L.on_device_of
def index_points_backward(g, idx, shape):
    """g [B,S,(K,)C] -> d points [B,N,C]: ordered segmented sum over a source-sorted CSR instead of
    index_put_(accumulate=True)'s float atomics."""
    B, N, Cc = shape
    dev = g.device
    M = idx[0].numel()
    keys = idx.reshape(B, M).to(torch.int32).contiguous()
    src = TTensor.from_rowmajor(g.reshape(B * M, Cc).contiguous())
    offs, perm = _csr(keys, B, M, N, dev)
    dst = TTensor(B * N, Cc, dev)
    L.psg_segment_sum(src.ptr, src.wchunks, 0, M, 1, None, offs.data_ptr(), perm.data_ptr(), M, N, B, Cc, dst.ptr,
                      dst.wchunks, 0, 0, _stream())
    return dst.to_rowmajor().reshape(B, N, Cc)

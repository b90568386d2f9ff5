"""Synthetic S3DIS-shaped inputs, labels and checkpoints.

There is no dataset and no trained checkpoint in the reference drop, so every measurement and every
parity test runs on data made here.  The generators are deterministic functions of a seed on the
torch CPU generator, so the build container (where the real reference runs and the golden vectors
are made) and the GPU box see identical tensors.

Block layout follows PointNet/data_utils/S3DISDataLoader.py:154-165 of the reference: channels
0:3 block-centred x,y and raw z (metres), 3:6 rgb/255, 6:9 xyz / room extent.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

NUM_CLASSES = 13

# (npoint, [radius...], [nsample...], [[mlp...]...]) per set-abstraction level; FP mlps; reference:
# PointNet/models/pointnet2_sem_seg.py:9-20 and pointnet2_sem_seg_msg.py:10-21.
ARCH = {
    "ssg": {
        "sa": [
            (1024, [0.1], [32], [[32, 32, 64]]),
            (256, [0.2], [32], [[64, 64, 128]]),
            (64, [0.4], [32], [[128, 128, 256]]),
            (16, [0.8], [32], [[256, 256, 512]]),
        ],
        "sa_in": [9, 64, 128, 256],
        "fp": [(768, [256, 256]), (384, [256, 256]), (320, [256, 128]), (128, [128, 128, 128])],
    },
    "msg": {
        "sa": [
            (1024, [0.05, 0.1], [16, 32], [[16, 16, 32], [32, 32, 64]]),
            (256, [0.1, 0.2], [16, 32], [[64, 64, 128], [64, 96, 128]]),
            (64, [0.2, 0.4], [16, 32], [[128, 196, 256], [128, 196, 256]]),
            (16, [0.4, 0.8], [16, 32], [[256, 256, 512], [256, 384, 512]]),
        ],
        "sa_in": [9, 96, 256, 512],
        "fp": [(1536, [256, 256]), (512, [256, 256]), (352, [256, 128]), (128, [128, 128, 128])],
    },
}


def make_blocks(B: int, N: int = 4096, seed: int = 0, kind: str = "uniform") -> torch.Tensor:
    """Return a [B, 9, N] float32 *non-contiguous* view (a transposed [B, N, 9] tensor), which is
    exactly how the reference scripts hand blocks to the model (NB_nontarget_test_semseg.py:163-165).

    kinds: uniform | grid | clustered | duplicates | surface
    """
    g = torch.Generator("cpu").manual_seed(seed)
    pts = torch.rand(B, N, 9, generator=g, dtype=torch.float32)
    if kind == "grid":
        # coordinates on multiples of 2^-6 so that every distance formula is exact in fp32; many ties
        pts[..., 0:3] = torch.floor(pts[..., 0:3] * 64.0) / 64.0
    elif kind == "clustered":
        # everything inside a 0.15 m cube: ball queries overflow nsample
        pts[..., 0:3] = pts[..., 0:3] * 0.15 + 0.4
    elif kind == "duplicates":
        # S3DIS blocks are padded by repeating points (S3DISDataLoader.py:149-153)
        half = N // 2
        src = torch.randint(0, half, (B, N - half), generator=g)
        for b in range(B):
            pts[b, half:] = pts[b, src[b]]
    elif kind == "surface":
        # points on five jittered planes: realistic local density
        plane = torch.randint(0, 5, (B, N), generator=g)
        jitter = (torch.rand(B, N, generator=g) - 0.5) * 0.01
        level = plane.float() * 0.2 + 0.05 + jitter
        axis = plane % 3
        for a in range(3):
            sel = axis == a
            col = pts[..., a]
            col[sel] = level[sel]
    elif kind != "uniform":
        raise ValueError(f"unknown synthetic kind {kind!r}")
    pts[..., 0:2] -= 0.5
    pts[..., 2] *= 3.0
    return pts.transpose(2, 1)


def zband_labels(x: torch.Tensor) -> torch.Tensor:
    """Spatially coherent labels for the targeted attacks: class = min(12, floor(13 z / 3))."""
    z = x[:, 2, :]
    return torch.clamp(torch.floor(z * (13.0 / 3.0)), 0, NUM_CLASSES - 1).to(torch.int64)


def class_palette() -> torch.Tensor:
    """[13, 3] fixed class colours in [0.15, 0.85] of the *painted* synthetic blocks."""
    g = torch.Generator("cpu").manual_seed(77)
    return 0.15 + 0.7 * torch.rand(NUM_CLASSES, 3, generator=g, dtype=torch.float32)


def make_painted_blocks(B: int, N: int = 4096, seed: int = 0, noise: float = 0.08):
    """Synthetic *labelled* blocks for training and for informative attack metrics: uniform blocks
    (``make_blocks``) cut into 13 horizontal slabs whose class is painted into the colour channels, colour =
    palette[class] + noise * N(0,1) clipped to [0,1].  The slab pattern is shifted cyclically by a random offset per
    block (class = floor(13 * frac(z / 3 + u_b))), so the absolute height says nothing about the class: a network
    trained on these blocks reads the class from the colours of a neighbourhood, and a colour attack can move its
    predictions -- which a random-init network's metrics never show.
    Returns (x [B,9,N] non-contiguous view like make_blocks, labels int64 [B,N])."""
    x = make_blocks(B, N, seed, "uniform")
    g = torch.Generator("cpu").manual_seed(seed + 7919)
    shift = torch.rand(B, 1, generator=g, dtype=torch.float32)
    frac = torch.remainder(x[:, 2, :] / 3.0 + shift, 1.0)
    labels = torch.clamp(torch.floor(frac * NUM_CLASSES), 0, NUM_CLASSES - 1).to(torch.int64)
    col = class_palette()[labels] + noise * torch.randn(B, N, 3, generator=g, dtype=torch.float32)
    x[:, 3:6] = col.clamp_(0.0, 1.0).transpose(2, 1)
    return x, labels


def _conv_keys(prefix, cin, cout, ndim):
    shape = (cout, cin, 1, 1) if ndim == 2 else (cout, cin, 1)
    return [(prefix + ".weight", shape, cin), (prefix + ".bias", (cout,), cin)]


def state_dict_spec(arch: str):
    """(key, shape, fan_in | 'bn_*') list with the reference's checkpoint key names."""
    a = ARCH[arch]
    spec = []
    for li, (npoint, radii, nsamples, mlps) in enumerate(a["sa"]):
        cin0 = a["sa_in"][li] + 3
        for bi, mlp in enumerate(mlps):
            cin = cin0
            for j, cout in enumerate(mlp):
                if arch == "ssg":
                    cp, bp = f"sa{li+1}.mlp_convs.{j}", f"sa{li+1}.mlp_bns.{j}"
                else:
                    cp, bp = f"sa{li+1}.conv_blocks.{bi}.{j}", f"sa{li+1}.bn_blocks.{bi}.{j}"
                spec += _conv_keys(cp, cin, cout, 2)
                spec += _bn_keys(bp, cout)
                cin = cout
    for fi, (cin, mlp) in enumerate(a["fp"]):
        name = f"fp{4-fi}"
        for j, cout in enumerate(mlp):
            spec += _conv_keys(f"{name}.mlp_convs.{j}", cin, cout, 1)
            spec += _bn_keys(f"{name}.mlp_bns.{j}", cout)
            cin = cout
    spec += _conv_keys("conv1", 128, 128, 1)
    spec += _bn_keys("bn1", 128)
    spec += _conv_keys("conv2", 128, NUM_CLASSES, 1)
    return spec


def _bn_keys(prefix, c):
    return [
        (prefix + ".weight", (c,), "bn_w"),
        (prefix + ".bias", (c,), "bn_b"),
        (prefix + ".running_mean", (c,), "bn_m"),
        (prefix + ".running_var", (c,), "bn_v"),
        (prefix + ".num_batches_tracked", (), "bn_n"),
    ]


def make_state_dict(arch: str = "ssg", seed: int = 1234, randomize_bn: bool = True,
                    init: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """A random checkpoint with the reference's key names and shapes.

    ``init="default"`` (what the golden vectors were made with) uses PyTorch's default conv scale, under
    which the signal shrinks layer by layer and the random network predicts one class everywhere -- fine
    for numerics parity, useless for attack-success metrics.  ``init="he"`` draws conv weights from
    U(+-sqrt(6 / fan_in)) (variance-preserving through ReLU) with zero conv biases and keeps BatchNorm
    near identity, which gives a network whose per-point predictions depend on the input, so that the
    attacks visibly move accuracy / mIoU / target hit-rate.

    Conv weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default scale).  With
    ``randomize_bn`` the BatchNorm affine parameters and running statistics are randomised too
    (running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5), weight ~ U(0.5,1.5), bias ~ N(0,0.1));
    default-initialised BN (mean 0 / var 1) would hide folding bugs.
    """
    g = torch.Generator("cpu").manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(arch):
        if kind == "bn_n":
            sd[key] = torch.tensor(0, dtype=torch.int64)
        elif init == "he" and kind in ("bn_w", "bn_v", "bn_b", "bn_m"):
            r = torch.rand(shape, generator=g)
            sd[key] = {"bn_w": 0.9 + 0.2 * r, "bn_v": 0.9 + 0.2 * r, "bn_b": 0.02 * (r - 0.5), "bn_m": 0.02 * (r - 0.5)}[kind]
        elif init == "he" and key.endswith(".bias"):
            torch.rand(shape, generator=g)
            sd[key] = torch.zeros(shape)
        elif init == "he":
            sd[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * math.sqrt(6.0 / kind)
        elif kind == "bn_w":
            sd[key] = torch.rand(shape, generator=g) + 0.5 if randomize_bn else torch.ones(shape)
        elif kind == "bn_v":
            sd[key] = torch.rand(shape, generator=g) + 0.5 if randomize_bn else torch.ones(shape)
        elif kind in ("bn_b", "bn_m"):
            sd[key] = torch.randn(shape, generator=g) * 0.1 if randomize_bn else torch.zeros(shape)
        else:
            bound = 1.0 / math.sqrt(kind)
            sd[key] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
    return sd


def load_checkpoint(arch: str = "ssg") -> "OrderedDict[str, torch.Tensor]":
    """The synthetic *trained* checkpoint shipped with the tests (tests/golden/ckpt_<arch>_painted.npz: 300 Adam steps
    on ``make_painted_blocks`` from the ``init="he"`` network, made by oracle/make_checkpoint.py; stored rounded to
    float16, widened here exactly).  On it the colour attacks visibly move accuracy / mIoU / target hit-rate."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"ckpt_{arch}_painted.npz")
    z = np.load(path)
    sd = OrderedDict()
    for key, _, _ in state_dict_spec(arch):
        v = torch.from_numpy(z[key])
        sd[key] = v.float() if v.dtype == torch.float16 else v
    return sd


def make_room(P: int, seed: int = 0, kind: str = "box") -> np.ndarray:
    """Synthetic S3DIS-shaped room: float64 [P, 7] = (x, y, z in metres, r, g, b in 0..255, label 0..12), the layout
    of the reference's ``stanford_indoor3d/Area_*.npy`` files (data_utils/S3DISDataLoader.py:105-108).

    kinds: ``box``     uniform in a 3.3 x 2.4 x 2.8 m room, random point order;
           ``objects`` the same points stored object by object (sorted into 0.4 m tiles), as S3DIS rooms are;
           ``lshape``  an L-shaped 4.1 x 3.6 m room: one quadrant is empty (empty grid columns) and one strip is
                       thinly populated (columns that need sampling WITH replacement);
           ``tiny``    a room barely larger than one block with fewer points than one block holds.
    """
    rng = np.random.RandomState(1000 + seed)
    if kind in ("box", "objects"):
        xy = rng.rand(P, 2) * np.array([3.3, 2.4]) + np.array([-1.7, 0.4])
    elif kind == "lshape":
        xy = rng.rand(P, 2) * np.array([4.1, 3.6])
        dead = (xy[:, 0] > 2.3) & (xy[:, 1] > 2.0)
        xy[dead, 0] -= 2.3                                  # fold the empty quadrant back into the room
        thin = rng.rand(P) < 0.01
        xy[thin] = rng.rand(int(thin.sum()), 2) * np.array([0.55, 0.9]) + np.array([2.45, 2.6])   # a thin strip inside it
    elif kind == "tiny":
        xy = rng.rand(P, 2) * np.array([1.2, 1.1]) + np.array([5.0, -3.0])
    else:
        raise ValueError(kind)
    z = rng.rand(P) * 2.8
    rgb = np.floor(rng.rand(P, 3) * 256.0)
    label = np.minimum(12, np.floor(13.0 * z / 2.8))
    room = np.concatenate([xy, z[:, None], rgb, label[:, None]], axis=1).astype(np.float64)
    if kind == "objects":
        tile = np.floor(room[:, 0] / 0.4) * 100 + np.floor(room[:, 1] / 0.4)
        room = room[np.argsort(tile, kind="stable")]
    return np.ascontiguousarray(room)

"""ctypes front end of oracle/geom_oracle.c (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpsg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "geom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps(xyz: torch.Tensor, npoint: int, start: torch.Tensor) -> torch.Tensor:
    a = _f32(xyz)
    B, N, _ = a.shape
    st = np.ascontiguousarray(start.cpu().numpy(), dtype=np.int64)
    out = np.empty((B, npoint), dtype=np.int64)
    lib().oracle_fps(_p(a), _p(st), B, N, npoint, _p(out))
    return torch.from_numpy(out)


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    a, b = _f32(src), _f32(dst)
    B, N, _ = a.shape
    M = b.shape[1]
    out = np.empty((B, N, M), dtype=np.float32)
    lib().oracle_square_distance(_p(a), _p(b), B, N, M, _p(out))
    return torch.from_numpy(out)


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    a, q = _f32(xyz), _f32(new_xyz)
    B, N, _ = a.shape
    S = q.shape[1]
    out = np.empty((B, S, nsample), dtype=np.int64)
    lib().oracle_ball_query(_p(a), _p(q), B, N, S, ctypes.c_double(radius), nsample, _p(out))
    return torch.from_numpy(out)


def three_nn(xyz1: torch.Tensor, xyz2: torch.Tensor):
    a, b = _f32(xyz1), _f32(xyz2)
    B, N, _ = a.shape
    S = b.shape[1]
    idx = np.empty((B, N, 3), dtype=np.int64)
    d2 = np.empty((B, N, 3), dtype=np.float32)
    w = np.empty((B, N, 3), dtype=np.float32)
    lib().oracle_three_nn(_p(a), _p(b), B, N, S, _p(idx), _p(d2), _p(w))
    return torch.from_numpy(idx), torch.from_numpy(d2), torch.from_numpy(w)

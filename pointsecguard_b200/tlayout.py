"""T-layout activations on the Python side (see csrc/psg_common.cuh): float T[rows/128][C/4][128][4],
rows padded to 128, channels to 16.  Only a thin holder + the pack / unpack C-ABI calls."""
from __future__ import annotations

import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


class TTensor:
    def __init__(self, rows: int, channels: int, device, zero: bool = False):
        self.rows = rows
        self.channels = channels
        self.cpad = (channels + 15) // 16 * 16
        self.rpad = (rows + 127) // 128 * 128
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc(self.rpad * self.cpad, dtype=torch.float32, device=device)

    @property
    def ptr(self):
        return self.buf.data_ptr()

    @property
    def wchunks(self):
        return self.cpad // 4

    @staticmethod
    def from_channels_first(x: torch.Tensor, want_xyz: bool = False):
        """x [B,C,N] (any strides) -> T [B*N][C]; optionally also xyz = x[:, :3] as [B,N,3]."""
        B, C, N = x.shape
        t = TTensor(B * N, C, x.device)
        xyz = torch.empty(B, N, 3, dtype=torch.float32, device=x.device) if want_xyz else None
        sb, sc, sn = x.stride()
        L.psg_pack_channels_first(x.data_ptr(), sb, sc, sn, B, C, N, t.ptr, t.wchunks, 0,
                                  xyz.data_ptr() if want_xyz else None, _stream())
        return (t, xyz) if want_xyz else t

    @staticmethod
    def from_rowmajor(x: torch.Tensor):
        """x [rows, C] -> T [rows][C]."""
        return TTensor.from_channels_first(x.t().unsqueeze(0))

    def to_channels_first(self, B: int, N: int, channels: int | None = None, c0: int = 0) -> torch.Tensor:
        C = self.channels if channels is None else channels
        y = torch.empty(B, C, N, dtype=torch.float32, device=self.buf.device)
        L.psg_unpack_channels_first(self.ptr, self.wchunks, c0 // 4, B, C, N, y.data_ptr(), 0, _stream())
        return y

    def to_rowmajor(self) -> torch.Tensor:
        return self.to_channels_first(1, self.rows)[0].t().contiguous()

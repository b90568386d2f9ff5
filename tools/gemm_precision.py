#!/usr/bin/env python
"""Error of ONE shared-MLP layer (Y = relu(X W^T + b), rows x K -> N) in the three MLP modes against the float64 product:
CUDA-core fp32 (mode 0), tcgen05 TF32 with the truncation-compensated weights (mode 1), tcgen05 3xTF32 (mode 2).
    python tools/gemm_precision.py > profiles/r2_gemm_precision.txt"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointsecguard_b200 import _lib as L
from pointsecguard_b200.tlayout import TTensor

st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
print("rows x K -> N | max |err| / max |Y| and rms(err) / rms(Y) per mode (against float64)")
for rows, K, N in ((4096, 16, 32), (4096, 64, 64), (4096, 128, 128), (2048, 256, 256), (1024, 768, 256)):
    x = torch.randn(rows, K, device="cuda", generator=g)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    b = torch.zeros(N, device="cuda")
    ref = (x.double() @ w.double().t() + b.double())
    wn, bn = np.ascontiguousarray(w.cpu().numpy()), np.ascontiguousarray(b.cpu().numpy())
    h = L.psg_mlp_create(wn.ctypes.data, bn.ctypes.data, K, N)
    tx = TTensor.from_rowmajor(x)
    out = []
    for mode in (0, 1, 2):
        ty = TTensor(rows, N, x.device, zero=True)
        L.psg_mlp_forward(h, tx.ptr, tx.wchunks, 0, tx.wchunks, None, 0, 0, 0, rows, ty.ptr, ty.wchunks, 0, mode, st)
        y = ty.to_rowmajor().double()
        e = y - ref
        scale = ((e * ref).sum() / (ref * ref).sum()).item()          # the part of the error that is a pure scaling of Y
        resid = e - scale * ref
        out.append((e.abs().max().item() / ref.abs().max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(),
                    scale, (resid.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()))
    L.psg_mlp_destroy(h)
    print(f"{rows} x {K} -> {N}: fp32 {out[0][0]:.2e} / {out[0][1]:.2e}   tf32 {out[1][0]:.2e} / {out[1][1]:.2e}   3xtf32 {out[2][0]:.2e} / {out[2][1]:.2e}"
          f"   | scale part of the error and rms of the rest: fp32 {out[0][2]:+.2e} {out[0][3]:.2e}  tf32 {out[1][2]:+.2e} {out[1][3]:.2e}  3xtf32 {out[2][2]:+.2e} {out[2][3]:.2e}")

#!/usr/bin/env python
"""Informative second baseline: the reference's algorithm in STOCK PyTorch ops on the same B200 -- what a user
gets by moving the reference's own code to the GPU (Python FPS loop of 1360 iterations per forward, matmul + full sort for
ball query and 3-NN, cuDNN / cuBLAS convolutions, autograd, index_put atomics).  The reference's scripts cannot run as they
are (hard-coded dataset / checkpoint paths, numpy API removed since), so this drives the oracle's op-for-op torch
restatement (oracle/pointnet2_oracle.py with GEOMETRY = "torch", pinned to the reference on the CPU) with CUDA tensors.
NB_attack only (the NU attacks of the reference index a CPU tensor with a CUDA index and fail on a GPU).

    python tools/stock_torch_baseline.py [--blocks 16] [--iters 5] [--out profiles/r2_stock_torch_b200.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import attacks_oracle as AO, pointnet2_oracle as PO          # noqa: E402
from pointsecguard_b200 import synthetic as syn                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    PO.GEOMETRY = "torch"
    sd = {k: v.to(dev) for k, v in syn.make_state_dict("ssg", init="he").items()}
    model = PO.OracleModel(sd, "ssg")
    x, labels = syn.make_painted_blocks(args.blocks, 4096, 0)
    xd = x.to(dev)
    lab = labels.numpy().astype(np.float64)
    res = {}
    for tf32 in (False, True):
        # TF32 is only ever allowed for the cuDNN convolutions: with allow_tf32 matmuls, square_distance's expansion formula is
        # off by ~1e-2 m^2 at these coordinates, a centroid falls outside its own ball and query_ball_point indexes out of range
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(0)
        AO.nb_attack(model, xd, lab, eps=0.1, alpha=0.05, iters=1)          # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        adv = AO.nb_attack(model, xd, lab, eps=0.1, alpha=0.05, iters=args.iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["cudnn_tf32_convs" if tf32 else "fp32"] = {"steps_per_s": args.iters / dt, "ms_per_step": dt / args.iters * 1e3,
                                            "peak_memory_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
    # the same attack through this repository (TF32 mode and fp32 mode)
    from pointsecguard_b200 import torchattacks
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    m = get_model(13)
    m.load_state_dict(syn.make_state_dict("ssg", init="he"))
    m = m.to(dev).eval()
    for mode, name in ((MLP_FP32, "fp32"), (MLP_TF32, "tf32")):
        m.set_mlp_mode(mode)
        atk = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=50)
        atk(xd, lab)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        atk(xd, lab)
        torch.cuda.synchronize()
        res["this_repo_" + name] = {"steps_per_s": 50 / (time.perf_counter() - t0)}
    out = {"workload": f"SSG NB_attack eps=0.1 alpha=0.05, B={args.blocks} x 4096, random-init he network", "stock_torch": res,
           "note": "stock PyTorch = oracle/pointnet2_oracle.py (GEOMETRY torch: the reference's op-for-op path) on cuda:0"}
    print(json.dumps(out))
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

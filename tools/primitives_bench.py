#!/usr/bin/env python
"""Primitive microbenchmark + scale sweep (BASELINE.json configs[4]): FPS, ball query, 3-NN and
index gather at N = 4096 / 16384 / 65536 points per block against the HBM roofline (SURVEY.md 8d
byte formulas, measured copy bandwidth from MEASURED_PEAKS.json), and NB-attack steps/s at each N.

    python tools/primitives_bench.py [--blocks 8] [--iters 10] [--out profiles/r1_primitives.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from pointsecguard_b200 import synthetic as syn, torchattacks
    from pointsecguard_b200.engine import MLP_TF32
    from pointsecguard_b200.models import pointnet_util as PU
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    try:
        hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except Exception:
        hbm, src = 6650.0, "fallback"
    B = args.blocks
    model = get_model(13)
    model.load_state_dict(syn.make_state_dict("ssg"))
    model = model.cuda().eval()
    model.set_mlp_mode(MLP_TF32)
    out = {"blocks": B, "hbm_gbs": hbm, "hbm_source": src, "sizes": {}}
    for N in (4096, 16384, 65536):
        x = syn.make_blocks(B, N, 0).cuda()
        xyz = x[:, :3].permute(0, 2, 1).contiguous()
        S, K, r = 1024, 32, 0.1
        start = torch.randint(0, N, (B,))
        fps = torch.ops.psg.fps(xyz, S, start)
        new_xyz = PU.index_points(xyz, fps)
        idx = PU.query_ball_point(r, K, xyz, new_xyz)
        pts = x.permute(0, 2, 1).contiguous()
        res = {}
        t = timed(lambda: torch.ops.psg.fps(xyz, S, start))
        res["fps"] = {"ms": t, "bytes": B * (12 * N + 8 * S), "rounds_per_s": B * S / (t / 1e3)}
        t = timed(lambda: PU.query_ball_point(r, K, xyz, new_xyz))
        res["ball_query"] = {"ms": t, "bytes": B * (12 * N + 12 * S + 8 * S * K)}
        t = timed(lambda: torch.ops.psg.three_nn(xyz, new_xyz))
        res["three_nn"] = {"ms": t, "bytes": B * (12 * N + 12 * S + 3 * N * 12)}
        t = timed(lambda: PU.index_points(pts, idx))
        res["gather_C9"] = {"ms": t, "bytes": B * (8 * S * K + 4 * 9 * N + 4 * 9 * S * K)}
        for k, v in res.items():
            v["GBps"] = v["bytes"] / (v["ms"] / 1e3) / 1e9
            v["frac_of_hbm"] = v["GBps"] / hbm
        lab = syn.zband_labels(x.cpu()).numpy().astype(np.float64)
        atk = torchattacks.NB_attack(model, eps=0.1, alpha=0.05, iters=args.iters)
        t = timed(lambda: atk(x, lab), reps=3, warm=1)
        res["nb_attack"] = {"ms_per_step": t / args.iters, "steps_per_s": args.iters / (t / 1e3),
                            "block_steps_per_s": B * args.iters / (t / 1e3)}
        out["sizes"][str(N)] = res
        print(N, json.dumps(res), flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

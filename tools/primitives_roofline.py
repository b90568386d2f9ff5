#!/usr/bin/env python
"""HBM roofline of the geometric primitives at SATURATING sizes, timed straight through the C ABI (ctypes, caller-owned
buffers, int32 indices -- no Python-side widening or allocation inside the timed region).

north_star names FPS, ball query, gather and interpolation against the HBM roofline; at the attack's own batch sizes
these launches are tens of microseconds over a few MB (latency-bound), so this tool also runs them with enough problems
in flight to fill the GPU: P = 256 clouds of 4096 points, 64 of 16384, 16 of 65536.  Algorithmic bytes are SURVEY.md 8d's
per-block formulas (int32 indices on this boundary: 4 bytes per index where the API formula says 8); achieved GB/s =
bytes x P / median launch time (CUDA events, 20 launches after 3 warm-ups, inputs larger than L2 or L2 flushed), against
the measured copy bandwidth of MEASURED_PEAKS.json.

    python tools/primitives_roofline.py [--out profiles/r2_primitives_roofline.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from pointsecguard_b200 import _lib as L             # noqa: E402
from pointsecguard_b200 import synthetic as syn      # noqa: E402
from pointsecguard_b200.tlayout import TTensor       # noqa: E402


def timed(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for i in range(reps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    try:
        hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm, src = 6650.0, "fallback"
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"hbm_gbs": hbm, "hbm_source": src, "cases": {}}
    for P, N in ((16, 4096), (256, 4096), (64, 16384), (16, 65536)):
        S, K, r, D = 1024, 32, 0.1, 128
        x = syn.make_blocks(min(P, 16), N, 0)
        xyz = x[:, :3].permute(0, 2, 1).contiguous().repeat((P + 15) // 16, 1, 1)[:P].contiguous().to(dev)
        xyz += torch.rand(P, 1, 3, device=dev) * 1e-3                       # distinct clouds
        start = torch.randint(0, N, (P,), dtype=torch.int32).to(dev)
        fps = torch.empty(P, S, dtype=torch.int32, device=dev)
        new_xyz = torch.empty(P, S, 3, device=dev)
        wsb = L.psg_fps_workspace(P, N)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        res = {}
        f = lambda: L.psg_fps(xyz.data_ptr(), P, P, N, S, start.data_ptr(), fps.data_ptr(), new_xyz.data_ptr(), ws.data_ptr(), wsb, st)
        res["fps"] = (timed(f, flush), 12 * N + 4 * S + 12 * S)
        ball = torch.empty(P, S, K, dtype=torch.int32, device=dev)
        gws_b = L.psg_ball_grid_workspace(P, N)
        gws = torch.empty(gws_b, dtype=torch.uint8, device=dev)
        rr, kk = (C.c_double * 1)(r), (C.c_int * 1)(K)
        f = lambda: L.psg_ball_query_grid(xyz.data_ptr(), P, P, N, new_xyz.data_ptr(), S, 1, rr, kk, ball.data_ptr(), None,
                                          gws.data_ptr(), gws_b, st)
        res["ball_query_grid (build + query)"] = (timed(f, flush), 12 * N + 12 * S + 4 * S * K)
        if N <= 16384:
            f = lambda: L.psg_ball_query(xyz.data_ptr(), P, P, N, new_xyz.data_ptr(), S, 1, rr, kk, ball.data_ptr(), None, st)
            res["ball_query_scan"] = (timed(f, flush, reps=8), 12 * N + 12 * S + 4 * S * K)
        nn_i = torch.empty(P, N, 3, dtype=torch.int32, device=dev)
        nn_w = torch.empty(P, N, 3, device=dev)
        f = lambda: L.psg_three_nn(xyz.data_ptr(), P, P, N, new_xyz.data_ptr(), S, nn_i.data_ptr(), nn_w.data_ptr(), None, st)
        res["three_nn"] = (timed(f, flush), 12 * N + 12 * S + 3 * N * 8)
        # interpolation: [P*S, D] coarse features (T-layout) -> [P*N, D]
        coarse = TTensor(P * S, D, dev, zero=True)
        fine = TTensor(P * N, D, dev)
        f = lambda: L.psg_interpolate(coarse.ptr, coarse.wchunks, S, nn_i.data_ptr(), nn_w.data_ptr(), P, N, D // 4, fine.ptr,
                                      fine.wchunks, 0, st)
        res["interpolate_D128"] = (timed(f, flush), 24 * N + 4 * D * S + 4 * D * N)
        # grouping gather (the SA1 shape: 9 feature channels + centred xyz -> 16-wide T-layout rows)
        feats = TTensor(P * N, 16, dev, zero=True)
        grouped = TTensor(P * S * K, 16, dev)
        f = lambda: L.psg_group_points(feats.ptr, feats.wchunks, 9, xyz.data_ptr(), P, N, new_xyz.data_ptr(), ball.data_ptr(), P, S, K,
                                       grouped.ptr, 16, st)
        res["group_C9+xyz"] = (timed(f, flush), 4 * S * K + 4 * 9 * N + 12 * N + 12 * S + 4 * 16 * S * K)
        # index_points at the API (row-major, int64 indices)
        pts = torch.rand(P, N, 64, device=dev)
        idx64 = ball.reshape(P, S * K).to(torch.int64)
        gout = torch.empty(P, S * K, 64, device=dev)
        f = lambda: L.psg_index_points(pts.data_ptr(), idx64.data_ptr(), P, N, 64, S * K, gout.data_ptr(), st)
        res["index_points_C64"] = (timed(f, flush), 8 * S * K + 4 * 64 * N + 4 * 64 * S * K)
        # dense kNN graph (ResGCN's graph op, csrc/knn.cu): k = 16 over the cloud itself, xyz and 64-wide features
        if N <= 16384:
            knn_out = torch.empty(P, N, 16, dtype=torch.int64, device=dev)
            f = lambda: L.psg_dense_knn(xyz.data_ptr(), P, N, 3, 16, knn_out.data_ptr(), None, st)
            res["dense_knn_C3_k16"] = (timed(f, flush, reps=5), 12 * N + 8 * 16 * N)
            if P * N <= 16 * 16384:
                f64 = pts[:, :, :64].contiguous()
                f = lambda: L.psg_dense_knn(f64.data_ptr(), P, N, 64, 16, knn_out.data_ptr(), None, st)
                res["dense_knn_C64_k16"] = (timed(f, flush, reps=3), 256 * N + 8 * 16 * N)
        case = {}
        for k, ((med, best), nbytes) in res.items():
            gb = nbytes * P / (med / 1e3) / 1e9
            case[k] = {"ms_median": med, "ms_best": best, "algorithmic_bytes": nbytes * P, "GBps": gb, "frac_of_hbm": gb / hbm}
        case["fps"]["rounds_per_s"] = P * S / (res["fps"][0][0] / 1e3)
        out["cases"][f"P={P} x N={N}"] = case
        print(f"P={P} N={N}", json.dumps({k: (round(v["ms_median"], 4), round(v["frac_of_hbm"], 4)) for k, v in case.items()}), flush=True)
        del xyz, pts, gout, idx64, coarse, fine, feats, grouped
        torch.cuda.empty_cache()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

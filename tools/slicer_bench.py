#!/usr/bin/env python
"""Whole-scene block slicer: GPU (csrc/slicer.cu through pointsecguard_b200.data_utils) against the CPU oracle
(numpy restatement of the reference's ScannetDatasetWholeScene.__getitem__) on one synthetic S3DIS-sized room.

    python tools/slicer_bench.py [--points 1000000] [--reps 5] [--no-cpu] > gpurun_out/slicer_bench.json

Prints one JSON line: blocks/s end to end (device-resident room -> float32 blocks on the device, host draws and the
count read-back included), per-kernel CUDA-event times with achieved GB/s against the measured HBM peak, and the CPU
oracle's blocks/s on the same room (bounded: one call)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from pointsecguard_b200 import _lib as L
    from pointsecguard_b200 import synthetic as syn
    from pointsecguard_b200.data_utils import S3DISDataLoader as PD
    P, bp = args.points, 4096
    side = max(2.0, (P / 10000.0) ** 0.5)          # ~10k points per square metre, as S3DIS rooms
    room = syn.make_room(P, 7, "objects")
    room[:, 0] = (room[:, 0] + 1.7) / 3.3 * side
    room[:, 1] = (room[:, 1] - 0.4) / 2.4 * side
    d = tempfile.mkdtemp()
    np.save(os.path.join(d, "Area_5_synthetic_1.npy"), room)
    ds = PD.ScannetDatasetWholeScene(d + "/", block_points=bp)
    np.random.seed(0)
    ds.blocks_device(0)                              # warm-up
    torch.cuda.synchronize()
    times = []
    for r in range(args.reps):
        np.random.seed(r)
        t0 = time.perf_counter()
        d32, lab, w, idx = ds.blocks_device(0)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    nb = d32.shape[0]
    # per-kernel-family device time (library profiler: CUDA event pairs on the launching stream)
    ds.stage_events = []
    np.random.seed(0)
    t0 = time.perf_counter()
    d32, lab, w, idx = ds.blocks_device(0)
    torch.cuda.synchronize()
    stages = {name: a.elapsed_time(b) for name, a, b in ds.stage_events}
    ds.stage_events = None
    rows = nb * bp
    members = rows                                   # upper bound of the membership lists (padding rows repeat members)
    byt = 2 * P * 16 + 2 * 4 * members + rows * (56 + 36 + 24 + 4 + 4)
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    line = {
        "what": "whole-scene block slicer (S3DISDataLoader.py:124-175), one room",
        "points": P, "blocks": int(nb), "block_points": bp,
        "gpu_wall_s": float(np.median(times)), "gpu_blocks_per_s": nb / float(np.median(times)),
        "device_stage_ms": stages, "device_ms": sum(stages.values()),
        "algorithmic_bytes": byt, "achieved_gbs": byt / (sum(stages.values()) / 1e3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
        "note": "wall time is dominated by the host's numpy draws (choice + shuffle per column, required for generator parity)",
    }
    if not args.no_cpu:
        from oracle import scene_slicer_oracle as SO
        lw = SO.label_weights([room[:, 6]])
        np.random.seed(0)
        t0 = time.perf_counter()
        out = SO.slice_room(room, lw, bp)
        dt = time.perf_counter() - t0
        line["cpu_oracle_s"] = dt
        line["cpu_blocks_per_s"] = out[0].shape[0] / dt
        line["speedup"] = dt / float(np.median(times))
        line["identical_to_oracle"] = bool(np.array_equal(out[3], idx.cpu().numpy()))
    print(json.dumps(line))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's whole-scene block slicer.

Follows PointNet/data_utils/S3DISDataLoader.py:
  * label_weights()  -> :115-122 (histogram of all labels, float32, (max / w) ** (1/3))
  * slice_room()     -> :124-175 (ScannetDatasetWholeScene.__getitem__)
It works on the point INDICES literally as the reference does (np.where -> np.random.choice on the
index list -> np.random.shuffle of the index list -> fancy indexing), which is what makes it a check
of the product's position-based draws.  Only tests/, __graft_entry__.smoke() and bench legs that time
the CPU baseline may import this module; nothing under pointsecguard_b200/ does.

Pinned against the reference itself: oracle/make_golden.py runs the unmodified reference class on
synthetic rooms and stores its outputs under tests/golden/scene_slicer_*.npz
(tests/test_oracle_golden.py::test_scene_slicer_oracle_matches_reference).
"""
from __future__ import annotations

import numpy as np


def label_weights(label_arrays, ncls=13):
    """S3DISDataLoader.py:115-122."""
    acc = np.zeros(ncls)
    for seg in label_arrays:
        hist, _ = np.histogram(seg, range(ncls + 1))
        acc += hist
    w = acc.astype(np.float32)
    w = w / np.sum(w)
    return np.power(np.amax(w) / w, 1 / 3.0)


def slice_room(room, labelweights, block_points=4096, stride=0.5, block_size=1.0, padding=0.001):
    """room: float64 [P, >=7] (x, y, z, r, g, b, label).  Returns (data_room [nb,bp,9] f64, label_room [nb,bp] int64,
    sample_weight [nb,bp] f64, index_room [nb,bp] int64) and consumes np.random's global state like the reference."""
    pts = room[:, :6]
    lab = room[:, 6]
    lo, hi = np.amin(pts, axis=0)[:3], np.amax(pts, axis=0)[:3]                               # :128
    nx = int(np.ceil(float(hi[0] - lo[0] - block_size) / stride) + 1)                           # :130
    ny = int(np.ceil(float(hi[1] - lo[1] - block_size) / stride) + 1)                           # :131
    out_d, out_l, out_w, out_i = [], [], [], []
    for iy in range(ny):                                                                        # :134-135
        for ix in range(nx):
            sx = lo[0] + ix * stride                                                            # :136-141
            ex = min(sx + block_size, hi[0])
            sx = ex - block_size
            sy = lo[1] + iy * stride
            ey = min(sy + block_size, hi[1])
            sy = ey - block_size
            member = (pts[:, 0] >= sx - padding) & (pts[:, 0] <= ex + padding) & \
                     (pts[:, 1] >= sy - padding) & (pts[:, 1] <= ey + padding)                  # :142-144
            ids = np.where(member)[0]
            if ids.size == 0:                                                                   # :145-146
                continue
            nbatch = int(np.ceil(ids.size / block_points))                                      # :147-148
            total = int(nbatch * block_points)
            with_replacement = not (total - ids.size <= ids.size)                               # :149
            extra = np.random.choice(ids, total - ids.size, replace=with_replacement)           # :150
            ids = np.concatenate((ids, extra))                                                  # :151
            np.random.shuffle(ids)                                                              # :152
            blk = pts[ids, :]                                                                   # :153 (a copy)
            rel = np.zeros((total, 3))                                                          # :154-157
            rel[:, 0] = blk[:, 0] / hi[0]
            rel[:, 1] = blk[:, 1] / hi[1]
            rel[:, 2] = blk[:, 2] / hi[2]
            blk[:, 0] = blk[:, 0] - (sx + block_size / 2.0)                                     # :158-159
            blk[:, 1] = blk[:, 1] - (sy + block_size / 2.0)
            blk[:, 3:6] /= 255.0                                                                # :160
            out_d.append(np.concatenate((blk, rel), axis=1))                                    # :161
            cls = lab[ids].astype(int)                                                          # :162
            out_l.append(cls)
            out_w.append(labelweights[cls].astype(np.float64))                                  # :163, widened by :167's hstack
            out_i.append(ids)
    data = np.concatenate(out_d, axis=0).reshape((-1, block_points, 9))                          # :165-172
    return (data, np.concatenate(out_l).reshape((-1, block_points)),
            np.concatenate(out_w).reshape((-1, block_points)), np.concatenate(out_i).reshape((-1, block_points)))

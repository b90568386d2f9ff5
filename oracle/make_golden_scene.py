"""Generate tests/golden/scene_slicer.npz by executing the UNMODIFIED reference class
``ScannetDatasetWholeScene`` (PointNet/data_utils/S3DISDataLoader.py:83-178) on synthetic rooms.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_scene

Rooms regenerate from seeds (pointsecguard_b200.synthetic.make_room), so the fixture holds outputs only:
per case the full index / label arrays (small integers), the label weights, a CRC of the float64 block
tensor and of the sample weights, and every 53rd data row.
"""
from __future__ import annotations

import os
import sys
import tempfile
import warnings
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/PointNet"
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from pointsecguard_b200 import synthetic as syn          # noqa: E402

# (name, room kind, points, room seed, block_points, stride, block_size, padding, numpy seed)
CASES = [
    ("box", "box", 6000, 0, 256, 0.5, 1.0, 0.001, 11),
    ("objects", "objects", 5000, 1, 128, 0.5, 1.0, 0.001, 12),
    ("lshape", "lshape", 7000, 2, 256, 0.5, 1.0, 0.001, 13),
    ("tiny", "tiny", 90, 3, 128, 0.5, 1.0, 0.001, 14),
    ("stride", "box", 4000, 4, 512, 0.75, 1.5, 0.01, 15),
]


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def run_reference(room, bp, stride, block_size, padding, npseed):
    # the reference module is imported here only: tests import CASES from this file on boxes without /root/reference
    sys.path.insert(0, os.path.join(REF, "data_utils"))
    warnings.filterwarnings("ignore")
    import S3DISDataLoader as RD
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "Area_5_synthetic_1.npy"), room)
        ds = RD.ScannetDatasetWholeScene(d + "/", block_points=bp, split="test", test_area=5, stride=stride,
                                         block_size=block_size, padding=padding)
        np.random.seed(npseed)
        out = ds[0]
        tail = np.random.randint(0, 1 << 30)            # the generator state after the call is part of the contract
        return out, ds.labelweights, tail


def main():
    from oracle import scene_slicer_oracle as SO
    out = {}
    for name, kind, P, rseed, bp, stride, bs, pad, npseed in CASES:
        room = syn.make_room(P, rseed, kind)
        (data, label, smpw, index), lw, tail = run_reference(room, bp, stride, bs, pad, npseed)
        assert data.dtype == np.float64 and smpw.dtype == np.float64, (data.dtype, smpw.dtype)
        out[f"{name}_index"] = index.astype(np.int32)
        out[f"{name}_label"] = label.astype(np.int8)
        out[f"{name}_lw"] = np.asarray(lw)
        out[f"{name}_data_crc"] = crc(data)
        out[f"{name}_smpw_crc"] = crc(smpw)
        out[f"{name}_rows"] = data.reshape(-1, 9)[::53].copy()
        out[f"{name}_tail"] = np.int64(tail)
        out[f"{name}_dtypes"] = np.array([str(index.dtype), str(label.dtype)])
        # report how the restatement compares
        np.random.seed(npseed)
        o = SO.slice_room(room, SO.label_weights([room[:, 6]]), bp, stride, bs, pad)
        same = all(np.array_equal(a, b) for a, b in zip(o, (data, label, smpw, index)))
        print(f"{name}: blocks {data.shape[0]} x {bp}, oracle identical to the reference: {same}")
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "scene_slicer.npz"), **out)
    print("wrote tests/golden/scene_slicer.npz",
          os.path.getsize(os.path.join(REPO, "tests", "golden", "scene_slicer.npz")), "bytes")


if __name__ == "__main__":
    main()

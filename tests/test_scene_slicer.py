"""Whole-scene block slicer (reference PointNet/data_utils/S3DISDataLoader.py:83-178).

CPU: the numpy oracle against the outputs of the unmodified reference class (tests/golden/scene_slicer.npz,
made by oracle/make_golden_scene.py) and the product's host logic (grid doubles, position-based draws, label weights).
GPU: the CUDA slicer through the C ABI against the oracle and the goldens, bit for bit, including numpy's generator
state after the call; a size-independent property test at S3DIS room size.
"""
import os
import tempfile
import zlib

import numpy as np
import pytest
import torch

from oracle import scene_slicer_oracle as SO
from oracle.make_golden_scene import CASES
from pointsecguard_b200 import synthetic as syn
from pointsecguard_b200.data_utils import S3DISDataLoader as PD

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "scene_slicer.npz"))


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def check_against_golden(name, out, lw, tail):
    data, label, smpw, index = out
    assert data.dtype == np.float64 and smpw.dtype == np.float64
    assert str(index.dtype) == GOLD[f"{name}_dtypes"][0] and str(label.dtype) == GOLD[f"{name}_dtypes"][1]
    assert np.array_equal(index, GOLD[f"{name}_index"])
    assert np.array_equal(label, GOLD[f"{name}_label"])
    assert np.array_equal(np.asarray(lw), GOLD[f"{name}_lw"])
    assert np.array_equal(data.reshape(-1, 9)[::53], GOLD[f"{name}_rows"])
    assert crc(data) == GOLD[f"{name}_data_crc"] and crc(smpw) == GOLD[f"{name}_smpw_crc"]
    assert tail == int(GOLD[f"{name}_tail"])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_scene_slicer_oracle_matches_reference(case):
    name, kind, P, rseed, bp, stride, bs, pad, npseed = case
    room = syn.make_room(P, rseed, kind)
    lw = SO.label_weights([room[:, 6]])
    np.random.seed(npseed)
    out = SO.slice_room(room, lw, bp, stride, bs, pad)
    tail = np.random.randint(0, 1 << 30)
    check_against_golden(name, out, lw, tail)


def test_golden_cases_cover_the_edge_cases():
    """empty columns, sampling with replacement, several blocks per column, clamped last column"""
    seen = {"empty": False, "replace": False, "multi": False}
    for name, kind, P, rseed, bp, stride, bs, pad, npseed in CASES:
        room = syn.make_room(P, rseed, kind)
        lo, hi = room[:, :3].min(0), room[:, :3].max(0)
        bounds, _ = PD.grid_columns(lo, hi, bs, stride, pad)
        for b in bounds:
            n = int(((room[:, 0] >= b[0]) & (room[:, 0] <= b[1]) & (room[:, 1] >= b[2]) & (room[:, 1] <= b[3])).sum())
            seen["empty"] |= n == 0
            seen["replace"] |= 0 < n < bp / 2
            seen["multi"] |= n > bp
    assert all(seen.values()), seen


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_host_draws_reproduce_the_reference_index_lists(case):
    """draw_positions works on member COUNTS only; applied to the np.where lists it must give the reference's
    index_room and leave numpy's generator in the same state."""
    name, kind, P, rseed, bp, stride, bs, pad, npseed = case
    room = syn.make_room(P, rseed, kind)
    lo, hi = np.amin(room[:, :3], axis=0), np.amax(room[:, :3], axis=0)
    bounds, centre = PD.grid_columns(lo, hi, bs, stride, pad)
    members = [np.where((room[:, 0] >= b[0]) & (room[:, 0] <= b[1]) & (room[:, 1] >= b[2]) & (room[:, 1] <= b[3]))[0]
               for b in bounds]
    np.random.seed(npseed)
    parts, block_cell = PD.draw_positions([m.size for m in members], bp)
    tail = np.random.randint(0, 1 << 30)
    nonempty = [m for m in members if m.size]
    index = np.concatenate([m[p] for m, p in zip(nonempty, parts)]).reshape(-1, bp)
    assert np.array_equal(index, GOLD[f"{name}_index"])
    assert tail == int(GOLD[f"{name}_tail"])
    assert len(block_cell) == index.shape[0]
    # block -> column map: every row of a block lies inside its column's padded bounds
    for blk, c in enumerate(block_cell):
        xy = room[index[blk], :2]
        b = bounds[c]
        assert (xy[:, 0] >= b[0]).all() and (xy[:, 0] <= b[1]).all() and (xy[:, 1] >= b[2]).all() and (xy[:, 1] <= b[3]).all()
    assert np.array_equal(PD.label_weights([room[:, 6]]), GOLD[f"{name}_lw"])


def test_slicer_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "Area_5_x.npy"), syn.make_room(500, 0, "tiny"))
        with pytest.raises(RuntimeError):
            PD.ScannetDatasetWholeScene(d + "/", block_points=128)


# ------------------------------------------------------------------------------------------------ GPU
def _dataset(room, bp, stride, bs, pad):
    d = tempfile.mkdtemp()
    np.save(os.path.join(d, "Area_5_synthetic_1.npy"), room)
    np.save(os.path.join(d, "Area_1_other.npy"), syn.make_room(300, 99, "tiny"))     # a training-split room: must be ignored
    return PD.ScannetDatasetWholeScene(d + "/", block_points=bp, split="test", test_area=5, stride=stride, block_size=bs,
                                       padding=pad)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_slicer_matches_reference_goldens_and_oracle(case):
    name, kind, P, rseed, bp, stride, bs, pad, npseed = case
    room = syn.make_room(P, rseed, kind)
    ds = _dataset(room, bp, stride, bs, pad)
    assert len(ds) == 1 and ds.scene_points_num == [P]
    assert np.array_equal(ds.room_coord_min[0], room[:, :3].min(0)) and np.array_equal(ds.room_coord_max[0], room[:, :3].max(0))
    np.random.seed(npseed)
    out = ds[0]
    tail = np.random.randint(0, 1 << 30)
    check_against_golden(name, out, ds.labelweights, tail)
    np.random.seed(npseed)
    ref = SO.slice_room(room, SO.label_weights([room[:, 6]]), bp, stride, bs, pad)
    for a, b in zip(out, ref):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # the device-resident variant: same draws, float32 rounding of torch.Tensor(ndarray)
    np.random.seed(npseed)
    d32, lab, w, idx = ds.blocks_device(0)
    assert d32.is_cuda and d32.dtype == torch.float32
    assert torch.equal(d32.cpu(), torch.Tensor(ref[0]).float())
    assert np.array_equal(idx.cpu().numpy(), ref[3]) and np.array_equal(lab.cpu().numpy(), ref[1])
    assert np.array_equal(w.cpu().numpy(), ref[2])


@pytest.mark.gpu
def test_gpu_slicer_room_scale_properties():
    """S3DIS room size (1M points, 4096-point blocks): properties that need no oracle run -- every block row lies in
    its column, rows are a permutation-with-repeats of the column's members, every point of the room is covered."""
    P, bp = 1_000_000, 4096
    room = syn.make_room(P, 7, "objects")
    ds = _dataset(room, bp, 0.5, 1.0, 0.001)
    np.random.seed(3)
    d32, lab, w, idx = ds.blocks_device(0)
    idx = idx.cpu().numpy()
    d = d32.cpu().numpy().astype(np.float64)
    src = room[idx.reshape(-1)]
    # columns 2..8 are pure functions of the source point
    hi = room[:, :3].max(0)
    assert np.array_equal(d32.cpu().numpy().reshape(-1, 9)[:, 2], src[:, 2].astype(np.float32))
    assert np.array_equal(d32.cpu().numpy().reshape(-1, 9)[:, 3:6], (src[:, 3:6] / 255.0).astype(np.float32))
    assert np.array_equal(d32.cpu().numpy().reshape(-1, 9)[:, 6:9], (src[:, :3] / hi).astype(np.float32))
    # block-centred x, y within half a block (+ padding)
    assert np.abs(d[..., 0]).max() <= 0.5 + 0.001 + 1e-6 and np.abs(d[..., 1]).max() <= 0.5 + 0.001 + 1e-6
    assert np.unique(idx).size == P                                   # every point of the room appears in some block
    assert np.array_equal(lab.cpu().numpy().reshape(-1), room[idx.reshape(-1), 6].astype(int))
    # a full oracle run of ONE column: ascending member list == np.where
    bounds, _ = PD.grid_columns(room[:, :3].min(0), hi, 1.0, 0.5, 0.001)
    b = bounds[len(bounds) // 2]
    members = np.where((room[:, 0] >= b[0]) & (room[:, 0] <= b[1]) & (room[:, 1] >= b[2]) & (room[:, 1] <= b[3]))[0]
    blocks_of_col = [k for k in range(idx.shape[0]) if set(idx[k]) <= set(members)]
    assert len(blocks_of_col) >= int(np.ceil(members.size / bp))


def test_scatter_last_wins_matches_numpy_assignment_on_cpu():
    """scene_eval scatters the perturbed points back into the scene with numpy's semantics for repeated indices
    (the last row wins, NB_nontarget_test_semseg.py:175-176); the helper is device-agnostic torch, checked here on CPU."""
    from pointsecguard_b200.scene_eval import _scatter_last_wins
    rng = np.random.default_rng(1)
    idx = rng.integers(0, 300, 2500)
    rows = rng.random((2500, 6)).astype(np.float32)
    ref = np.full((400, 6), -1.0, np.float32)
    ref[idx] = rows
    dst = torch.full((400, 6), -1.0)
    _scatter_last_wins(dst, torch.from_numpy(idx), torch.from_numpy(rows))
    assert np.array_equal(dst.numpy(), ref)

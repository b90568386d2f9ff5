"""Whole-scene block slicer on the GPU.

Reference: PointNet/data_utils/S3DISDataLoader.py:83-178, ``ScannetDatasetWholeScene``: same class
name, constructor arguments, public attributes (``scene_points_list``, ``semantic_labels_list``,
``room_coord_min`` / ``room_coord_max``, ``labelweights``, ``scene_points_num``) and
``__getitem__(index) -> (data_room [nb,bp,9] f64, label_room [nb,bp] int, sample_weight [nb,bp] f64,
index_room [nb,bp] int)``, bit for bit and consuming numpy's global generator exactly as the
reference does (one ``choice`` and one ``shuffle`` per non-empty column, in grid order).

What runs where:
  * device (csrc/slicer.cu through the C ABI): bounding box, ordered membership lists of every
    column (the reference's ``np.where`` per column), gather + normalisation of every block row;
    the room is uploaded once and stays resident;
  * host: the grid arithmetic (a handful of float64 scalars per column, written with the reference's
    own expressions so the bounds are the same doubles) and the random draws, which depend only on
    the member COUNT of a column: ``choice(point_idxs, k)`` draws positions exactly as
    ``choice(len(point_idxs), k)`` does, and shuffling a list is applying the shuffle of ``arange``.

``blocks_device(index)`` is the GPU-resident variant the attack pipeline uses: float32 ``[nb,bp,9]``
(what ``torch.Tensor(batch_data).float().cuda()`` yields in the scripts, NB_nontarget_test_semseg.py:
163-165), labels, weights and point indices as CUDA tensors -- no host round trip of the blocks.

There is no CPU fallback: the class raises without CUDA.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def draw_positions(totals, block_points):
    """S3DISDataLoader.py:146-152 on POSITIONS: for every non-empty column (grid order) pad its member list to a
    multiple of block_points with ``np.random.choice`` and shuffle it.  ``choice(point_idxs, k)`` picks positions exactly
    as ``choice(len(point_idxs), k)`` does and ``shuffle`` permutes by position, so drawing on ``arange(n)`` consumes
    numpy's global generator identically and ``point_idxs[pos]`` is the reference's shuffled index list.
    Returns (list of int64 position arrays, one per non-empty column; list with the column of every block)."""
    parts, block_cell = [], []
    for c, n in enumerate(int(t) for t in totals):
        if n == 0:
            continue
        num_batch = int(np.ceil(n / block_points))
        point_size = int(num_batch * block_points)
        replace = (point_size - n) > n
        extra = np.random.choice(n, point_size - n, replace=replace)
        pos = np.concatenate((np.arange(n), extra))
        np.random.shuffle(pos)
        parts.append(pos)
        block_cell.extend([c] * num_batch)
    return parts, block_cell


def grid_columns(coord_min, coord_max, block_size, stride, padding):
    """S3DISDataLoader.py:130-145: the grid_y x grid_x columns in loop order, written with the reference's own float64
    expressions so that the device compares against the very same doubles.  Returns (bounds [ncell,4] = padded
    (lo_x, hi_x, lo_y, hi_y), centre [ncell,2] = (s_x + block_size / 2, s_y + block_size / 2))."""
    grid_x = int(np.ceil(float(coord_max[0] - coord_min[0] - block_size) / stride) + 1)
    grid_y = int(np.ceil(float(coord_max[1] - coord_min[1] - block_size) / stride) + 1)
    bounds, centre = [], []
    for index_y in range(0, grid_y):
        s_y = coord_min[1] + index_y * stride
        e_y = min(s_y + block_size, coord_max[1])
        s_y = e_y - block_size
        for index_x in range(0, grid_x):
            s_x = coord_min[0] + index_x * stride
            e_x = min(s_x + block_size, coord_max[0])
            s_x = e_x - block_size
            bounds.append((s_x - padding, e_x + padding, s_y - padding, e_y + padding))
            centre.append((s_x + block_size / 2.0, s_y + block_size / 2.0))
    return np.asarray(bounds, dtype=np.float64).reshape(-1, 4), np.asarray(centre, dtype=np.float64).reshape(-1, 2)


def label_weights(label_arrays, ncls=13):
    """S3DISDataLoader.py:115-122: inverse-frequency class weights, float32, cube root."""
    hist = np.zeros(ncls)
    for seg in label_arrays:
        hist += np.histogram(seg, range(ncls + 1))[0]
    freq = hist.astype(np.float32)
    freq = freq / np.sum(freq)
    return np.power(np.amax(freq) / freq, 1 / 3.0)


class ScannetDatasetWholeScene:
    # S3DISDataLoader.py:84
    def __init__(self, root, block_points=4096, split='test', test_area=5, stride=0.5, block_size=1.0, padding=0.001,
                 device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ScannetDatasetWholeScene (B200 slicer) needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.block_points = block_points
        self.block_size = block_size
        self.padding = padding
        self.root = root
        self.split = split
        self.stride = stride
        self.scene_points_num = []
        assert split in ['train', 'test']
        tag = 'Area_%d' % test_area                    # :92-95: the test split is the rooms of one area
        self.file_list = [d for d in os.listdir(root) if (tag in d) == (split == 'test')]
        self.scene_points_list = []
        self.semantic_labels_list = []
        self.room_coord_min, self.room_coord_max = [], []
        self._rooms = []                       # device copies [P,7] float64, uploaded once
        self.stage_events = None               # set to [] to collect (name, start, end) CUDA events of the device stages
        ws = torch.zeros(L.psg_scene_minmax_workspace(), dtype=torch.uint8, device=self.device)
        self._ws = ws
        for file in self.file_list:
            data = np.load(root + file)
            self.scene_points_list.append(data[:, :6])
            self.semantic_labels_list.append(data[:, 6])
            room = torch.from_numpy(np.ascontiguousarray(data[:, :7], dtype=np.float64)).to(self.device)
            self._rooms.append(room)
            mm = self._minmax(room)
            self.room_coord_min.append(mm[:3].copy()), self.room_coord_max.append(mm[3:].copy())
        assert len(self.scene_points_list) == len(self.semantic_labels_list)

        self.scene_points_num = [seg.shape[0] for seg in self.semantic_labels_list]
        self.labelweights = label_weights(self.semantic_labels_list)
        self._lw_dev = torch.from_numpy(np.ascontiguousarray(self.labelweights, dtype=np.float32)).to(self.device)

    def _minmax(self, room):
        out = torch.empty(6, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            L.psg_scene_minmax(room.data_ptr(), room.shape[0], room.shape[1], out.data_ptr(), self._ws.data_ptr(),
                               self._ws.numel(), _stream())
        return out.cpu().numpy()

    def _timed(self, name, fn):
        if self.stage_events is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.stage_events.append((name, e0, e1))
        return out

    def _slice(self, index, want_f64, want_f32):
        room = self._rooms[index]
        P, ld = room.shape
        dev = self.device
        with torch.cuda.device(dev):
            st = _stream()
            mm = self._timed("minmax", lambda: self._minmax(room))     # :128 (recomputed per call, as the reference does)
            coord_min, coord_max = mm[:3], mm[3:]
            bounds, centre = grid_columns(coord_min, coord_max, self.block_size, self.stride, self.padding)
            ncell = bounds.shape[0]
            if ncell == 0:
                raise ValueError("room smaller than one block: the reference produces no blocks here")
            d_bounds = torch.from_numpy(bounds).to(dev)
            d_centre = torch.from_numpy(centre).to(dev)
            d_max = torch.from_numpy(np.ascontiguousarray(coord_max)).to(dev)
            nchunk = L.psg_scene_chunks(P)
            counts = torch.empty(ncell * nchunk, dtype=torch.int32, device=dev)
            totals = torch.empty(ncell, dtype=torch.int32, device=dev)
            self._timed("cell_counts", lambda: L.psg_scene_cell_counts(room.data_ptr(), P, ld, d_bounds.data_ptr(), ncell,
                                                                        counts.data_ptr(), totals.data_ptr(), st))
            tot = totals.cpu().numpy().astype(np.int64)              # the only sync: the draws depend on these counts
            pos_parts, block_cell = draw_positions(tot, self.block_points)
            bp = self.block_points
            row_pos = np.concatenate(pos_parts).astype(np.int32)
            rows = row_pos.shape[0]
            off = np.zeros(ncell, dtype=np.int64)
            off[1:] = np.cumsum(tot)[:-1]
            d_off = torch.from_numpy(off).to(dev)
            sel = torch.empty(int(tot.sum()), dtype=torch.int32, device=dev)
            self._timed("cell_fill", lambda: L.psg_scene_cell_fill(room.data_ptr(), P, ld, d_bounds.data_ptr(), ncell,
                                                                    counts.data_ptr(), d_off.data_ptr(), sel.data_ptr(), st))
            d_pos = torch.from_numpy(row_pos).to(dev)
            d_bc = torch.tensor(block_cell, dtype=torch.int32, device=dev)
            nb = rows // bp
            data = torch.empty(nb, bp, 9, dtype=torch.float64, device=dev) if want_f64 else None
            data32 = torch.empty(nb, bp, 9, dtype=torch.float32, device=dev) if want_f32 else None
            label = torch.empty(nb, bp, dtype=torch.int64, device=dev)
            smpw = torch.empty(nb, bp, dtype=torch.float64, device=dev)
            idx = torch.empty(nb, bp, dtype=torch.int64, device=dev)
            self._timed("gather", lambda: L.psg_scene_gather(room.data_ptr(), ld, 6, sel.data_ptr(), d_off.data_ptr(), d_bc.data_ptr(), d_pos.data_ptr(),
                               d_centre.data_ptr(), d_max.data_ptr(), self._lw_dev.data_ptr(), 13, rows, bp,
                               data.data_ptr() if data is not None else None,
                               data32.data_ptr() if data32 is not None else None,
                               label.data_ptr(), smpw.data_ptr(), idx.data_ptr(), st))
        return data, data32, label, smpw, idx

    def __getitem__(self, index):
        data, _, label, smpw, idx = self._slice(index, True, False)
        return data.cpu().numpy(), label.cpu().numpy(), smpw.cpu().numpy(), idx.cpu().numpy()

    def blocks_device(self, index):
        """-> (data float32 [nb,bp,9], label int64 [nb,bp], sample_weight float64 [nb,bp], point index int64 [nb,bp]),
        all CUDA tensors; draws from numpy's global generator exactly like ``__getitem__``."""
        _, data32, label, smpw, idx = self._slice(index, False, True)
        return data32, label, smpw, idx

    def __len__(self):
        return len(self.scene_points_list)

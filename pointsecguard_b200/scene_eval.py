"""Whole-scene attack evaluation: the per-scene loop of the attack scripts with every stage on the GPU.

Reference: PointNet/NB_nontarget_test_semseg.py:138-241 (and its NB_target / NU siblings): slice the room into
blocks (``TEST_DATASET_WHOLE_SCENE[batch_idx]``), per batch of blocks run the clean forward, the attack and the
adversarial forward, scatter the perturbed points back into the scene, vote the predictions into the scene pools,
count per-class seen / correct / union for blocks and for the voted scene.

Here: ``ScannetDatasetWholeScene.blocks_device`` (csrc/slicer.cu) -> ``get_model`` / ``torchattacks`` (the hot
path) -> ``metrics.VotePool`` / ``attack_counters`` (psg_add_vote, psg_confusion_matrix).  Nothing returns to the
host inside the loop; the draws on numpy's and torch's CPU generators happen in the reference's order (slicer, then
per batch: clean forward, attack forwards, adversarial forward), so a seeded run matches the reference's.
Multi-GPU (``evaluate_dataset``): scenes are independent, so rank r evaluates scenes r, r + W, ...; the only
exchange is one all-reduce of the global per-class counters (:272-291).  With ``seed`` given, both generators are
re-seeded per scene, which makes the W-GPU result equal the 1-GPU result scene for scene.
"""
from __future__ import annotations

import torch

from . import distributed as D
from . import metrics as MT


def _scatter_last_wins(dst, index, rows):
    """``dst[index] = rows`` with numpy's semantics for repeated indices (the LAST row wins), deterministically:
    an index_put with duplicates is unordered on the GPU, so the winning row of every target is found first."""
    n = index.numel()
    rid = torch.arange(n, device=index.device, dtype=torch.int64)
    last = torch.full((dst.shape[0],), -1, dtype=torch.int64, device=index.device)
    last.scatter_reduce_(0, index, rid, reduce="amax", include_self=True)
    hit = last >= 0
    dst[hit] = rows[last[hit]].to(dst.dtype)


def evaluate_scene(model, dataset, index, make_attack, batch_size=16, num_votes=1, num_classes=13):
    """-> dict with ``block`` / ``adv_block`` (MT.summarize of the per-block counters, :187-212), ``scene`` /
    ``adv_scene`` (MT.scene_iou of the voted pools, :216-241), ``adv_whole_scene`` ([P,6] float32 CUDA tensor, :176),
    ``dist`` (sum of the per-batch ``torch.dist(adv, x)`` of :177) and the raw counters / pools.

    ``make_attack()`` returns the attack object for one batch, e.g.
    ``lambda: torchattacks.NB_attack(model, eps=0.1, alpha=0.05, iters=10)`` (:169)."""
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("evaluate_scene needs the model on a CUDA device; there is no CPU fallback")
    P = dataset.scene_points_list[index].shape[0]
    pool, adv_pool = MT.VotePool(P, num_classes, device=dev), MT.VotePool(P, num_classes, device=dev)
    adv_whole_scene = torch.zeros(P, 6, dtype=torch.float32, device=dev)
    ncnt = num_classes * num_classes + 4
    cnt = torch.zeros(ncnt, dtype=torch.int64, device=dev)
    adv_cnt = torch.zeros(ncnt, dtype=torch.int64, device=dev)
    dist = torch.zeros((), dtype=torch.float64, device=dev)
    for _ in range(num_votes):
        data, label, smpw, pidx = dataset.blocks_device(index)           # [nb,bp,9] f32, [nb,bp] ...
        nb = data.shape[0]
        for s in range(0, nb, batch_size):
            e = min(s + batch_size, nb)
            x = data[s:e].transpose(2, 1)                                # [b,9,N] view, as :165
            lab = label[s:e]
            seg_pred, _ = model(x)                                       # :167
            adv = make_attack()(x, lab).detach()                         # :169-171
            adv_pred, _ = model(adv)                                     # :173
            _scatter_last_wins(adv_whole_scene, pidx[s:e].reshape(-1), adv.transpose(1, 2)[:, :, :6].reshape(-1, 6))   # :175-176
            dist += torch.dist(adv, x).double()                          # :177
            w = smpw[s:e].float()
            pool.add(seg_pred, pidx[s:e], w)                             # :180-185
            adv_pool.add(adv_pred, pidx[s:e], w)
            MT.attack_counters(seg_pred, lab, out=cnt)                   # :187-211
            MT.attack_counters(adv_pred, lab, out=adv_cnt)
    scene_label = torch.as_tensor(dataset.semantic_labels_list[index]).to(device=dev, dtype=torch.int32)
    scene_cnt, adv_scene_cnt = pool.counters(scene_label), adv_pool.counters(scene_label)
    return {
        "block": MT.summarize(cnt.cpu(), num_classes), "adv_block": MT.summarize(adv_cnt.cpu(), num_classes),
        "scene": MT.scene_iou(scene_cnt, num_classes), "adv_scene": MT.scene_iou(adv_scene_cnt, num_classes),
        "adv_whole_scene": adv_whole_scene, "dist": float(dist.item()),
        "counters": cnt, "adv_counters": adv_cnt, "pool": pool, "adv_pool": adv_pool,
    }


def evaluate_dataset(model, dataset, make_attack, batch_size=16, num_votes=1, num_classes=13, seed=None):
    """All scenes of ``dataset`` (the scripts' outer loop, :124-291), scene-strided over the ranks of the process group.
    -> dict with the global clean / adversarial scene counters (all-reduced), their IoU summaries, and this rank's
    per-scene results."""
    import numpy as np
    dev = next(model.parameters()).device
    ncnt = num_classes * num_classes + 4
    tot = torch.zeros(2, ncnt, dtype=torch.int64, device=dev)
    mine = {}
    for index in range(D.rank(), len(dataset), D.world_size()):
        if seed is not None:
            np.random.seed(seed + index)
            torch.manual_seed(seed + index)
        r = evaluate_scene(model, dataset, index, make_attack, batch_size, num_votes, num_classes)
        scene_label = torch.as_tensor(dataset.semantic_labels_list[index]).to(device=dev, dtype=torch.int32)
        r["pool"].counters(scene_label, out=tot[0])
        r["adv_pool"].counters(scene_label, out=tot[1])
        mine[index] = r
    D.all_reduce_sum_(tot)
    return {"counters": tot, "scene": MT.scene_iou(tot[0], num_classes), "adv_scene": MT.scene_iou(tot[1], num_classes),
            "scenes": mine}
